"""World-size-2 CPU tests (gloo) of the N>1 path's host logic and formulation.

What runs on the GPUs for N>1 (csrc/dist.cu) cannot run here; what can is
  * the launch-side plumbing of sprsolve_b200/dist.py (rendezvous, communicator-id distribution,
    row-block partition through the C ABI, max-over-ranks),
  * the FORMULATION of the partitioned solve: every rank owns a contiguous row block, a SpMV needs
    the neighbour planes of x (halo), every reduction is a local partial sum followed by a sum over
    ranks.  Two gloo ranks run exactly that with the oracle's kernels on their row blocks and must
    reproduce the serial oracle's Jacobi-BiCGStab residual history within the north star's 1e-10
    (only the summation order differs) -- the claim DESIGN.md section 5 makes for the GPU path.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 12  # grid of the 27-point system: 12^3 = 1728 rows, 2 slabs of 6 planes


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch
        import torch.distributed as dist

        import sprsolve_b200 as sp
        from oracle import oracle as orc
        from sprsolve_b200 import dist as spd

        torch.set_num_threads(1)
        w, r = spd.init_process_group("gloo")
        assert (w, r) == (world, rank) and spd.env_world() == (world, rank, rank)
        res = {}

        # ---- communicator-id distribution (rank 0 creates, everyone receives the same 128 bytes)
        class FakeCtx:
            got = None

            def comm_init(self, world_, rank_, uid):
                self.got = (world_, rank_, bytes(uid))

        ctx = FakeCtx()
        spd.attach_communicator(ctx, make_id=lambda: bytes(range(128)))
        assert ctx.got == (world, rank, bytes(range(128)))
        gathered = spd.all_gather_bytes(bytes([rank]) * 3)
        assert gathered == [bytes([q]) * 3 for q in range(world)]
        assert spd.max_over_ranks(10.0 + rank) == 10.0 + world - 1

        # ---- row-block partition through the C ABI
        rb, re = spd.row_block(sp.STENCIL_CONVDIFF27, G, G, G, world, rank)
        res["block"] = (rb, re)
        n = G**3
        nl = re - rb
        plane = G * G

        # ---- partitioned Jacobi-BiCGStab with the oracle's kernels on this rank's rows
        A = orc.gen_convdiff27(G, G, G, row_begin=rb, row_end=re)  # local rows, GLOBAL column ids
        lo_nb, hi_nb = rank - 1, rank + 1

        def spmv_dist(x_loc):
            xg = np.zeros(n)
            xg[rb:re] = x_loc
            reqs = []
            recv_lo, recv_hi = np.empty(plane), np.empty(plane)
            if lo_nb >= 0:
                reqs.append(dist.isend(torch.from_numpy(x_loc[:plane].copy()), lo_nb))
                reqs.append(dist.irecv(torch.from_numpy(recv_lo), lo_nb))
            if hi_nb < world:
                reqs.append(dist.isend(torch.from_numpy(x_loc[-plane:].copy()), hi_nb))
                reqs.append(dist.irecv(torch.from_numpy(recv_hi), hi_nb))
            for q in reqs:
                q.wait()
            if lo_nb >= 0:
                xg[rb - plane:rb] = recv_lo
            if hi_nb < world:
                xg[re:re + plane] = recv_hi
            return orc.spmv(A, xg)

        def gsum(*vals):
            t = torch.tensor(vals, dtype=torch.float64)
            dist.all_reduce(t)
            return [float(v) for v in t]

        ldot = lambda a, b: float(orc.conj_dot(a, b))  # noqa: E731  sequential fold (vecalg.rs:564-568)
        diag = np.full(nl, 27.75)
        dinv = 1.0 / diag
        rhs = spmv_dist(np.ones(nl))
        x = np.zeros(nl)
        (bb,) = gsum(ldot(rhs, rhs))
        rhs_norm = np.sqrt(bb)
        tol2 = 1e-8 * rhs_norm
        r = spmv_dist(x) - rhs  # r = A x - b (bicg_stab.rs:243-246)
        r0 = r.copy()
        (rr,) = gsum(ldot(r, r))
        hist = [np.sqrt(rr) / rhs_norm]
        rho = rr
        # unrolled iteration 0 (bicg_stab.rs:260-293)
        p = r.copy()
        y = p * dinv
        v = spmv_dist(y)
        (r0v,) = gsum(ldot(r0, v))
        alpha = rho / r0v
        r = r + v * (-alpha)
        z = r * dinv
        t = spmv_dist(z)
        tt, tr = gsum(ldot(t, t), ldot(t, r))
        wq = tr / tt if tt > 0 else 0.0
        x = x + y * (-alpha)
        x = x + z * (-wq)
        r = r + t * (-wq)
        its_done = None
        for its in range(1, 400):
            rr, rho_new = gsum(ldot(r, r), ldot(r0, r))
            rn = np.sqrt(rr)
            hist.append(rn / rhs_norm)
            if rn <= tol2:
                its_done = its
                break
            rho_old, rho = rho, rho_new
            beta = (rho / rho_old) * (alpha / wq)
            p = v * (-beta * wq) + p * beta
            p = p + r * 1.0
            y = p * dinv
            v = spmv_dist(y)
            (r0v,) = gsum(ldot(r0, v))
            alpha = rho / r0v
            r = r + v * (-alpha)
            z = r * dinv
            t = spmv_dist(z)
            tt, tr = gsum(ldot(t, t), ldot(t, r))
            wq = tr / tt if tt > 0 else 0.0
            x = x + y * (-alpha)
            x = x + z * (-wq)
            r = r + t * (-wq)
        res["hist"] = hist
        res["its"] = its_done
        res["x_err"] = float(np.abs(x - 1.0).max())
        out.put((rank, res))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback

        out.put((rank, {"error": f"{e}\n{traceback.format_exc()}"}))


def test_world_size_2_gloo(orc):
    world = 2
    mpctx = mp.get_context("spawn")
    out = mpctx.Queue()
    port = _free_port()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert "error" not in got[r], got[r]["error"]
    # partition: disjoint, ordered, plane-aligned cover of all rows
    b0, b1 = got[0]["block"], got[1]["block"]
    assert b0 == (0, G * G * (G // 2)) and b1 == (b0[1], G**3)
    # both ranks agree on every scalar of the recurrence
    assert got[0]["hist"] == got[1]["hist"] and got[0]["its"] == got[1]["its"]
    # the partitioned formulation reproduces the serial oracle (north star: 1e-10 over 50 iterations, +-2 %)
    A = orc.gen_convdiff27(G)
    rhs = orc.spmv(A, np.ones(A.n))
    o = orc.bicgstab(A, rhs, max_iter=400, tol=1e-8, pc=("diag", A.diagonal()), hist_cap=401)
    assert o.status == orc.OK
    h = np.array(got[0]["hist"])
    m = min(50, len(h), len(o.hist))
    live = o.hist[:m] >= 1e-8  # the terminal entry is only "< tol" (same rule as tests/test_gpu_parity.py)
    dev = np.where(live, np.abs(h[:m] - o.hist[:m]) / np.abs(o.hist[:m]), 0.0)
    assert np.all(dev <= 1e-10), dev
    assert abs(got[0]["its"] - o.iters) <= max(1, int(np.ceil(0.02 * o.iters)))
    assert max(got[0]["x_err"], got[1]["x_err"]) < 1e-6


def test_partition_properties():
    """spb_stencil_partition is host arithmetic: no GPU needed; blocks tile [0, n) at plane
    boundaries for every world size the bench uses, including uneven splits."""
    import sprsolve_b200 as sp
    from sprsolve_b200 import dist as spd

    for kind, (nx, ny, nz), plane in ((sp.STENCIL_CONVDIFF27, (8, 6, 10), 48), (sp.STENCIL_LAP3D7, (5, 5, 7), 25), (sp.STENCIL_DIRICHLET2D, (9, 9, 1), 9)):
        n = nx * ny * nz
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                b, e = spd.row_block(kind, nx, ny, nz, world, rank)
                assert b == prev and e >= b and b % plane == 0 and e % plane == 0
                prev = e
            assert prev == n
    with pytest.raises(ValueError):
        spd.row_block(sp.STENCIL_LAP3D7, 4, 4, 4, 2, 2)
