"""Multi-GPU parity (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).

One process per GPU, matrix partitioned by row blocks.  For both data-path transports (stores into
peer-mapped windows over NVLink -- the default -- and NCCL) and for both ways of building the
partitioned matrix (host CSR row blocks through spb_csr_create, on-device generator), checks
against the serial CPU oracle:
  * SpMV: bit-exact (each row is still the sequential CSR-order fold of src/mat.rs:100-105; halo
    columns are only renumbered),
  * mul_vec_dot: 1e-13 relative (summation order),
  * Jacobi-BiCGStab on the 27-point system: residual history within 1e-10 over the first 50
    iterations, iteration count +-2 %, solution within the tolerance,
  * unpreconditioned MINRES on the shifted 7-point system (strict bar as well).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _worker(rank, world, uid, transport, out):
    try:
        sys.path.insert(0, ROOT)
        os.environ["SPB_COMM"] = transport
        import sprsolve_b200 as sp
        from oracle import oracle as orc

        ctx = sp.Context(rank)
        ctx.comm_init(world, rank, uid)
        res = {}
        G = 24
        n = G**3
        # ---- (a) host CSR row blocks, global column ids
        rb, re = __import__("sprsolve_b200.dist", fromlist=["row_block"]).row_block(sp.STENCIL_CONVDIFF27, G, G, G, world, rank)
        Ab = orc.gen_convdiff27(G, G, G, row_begin=rb, row_end=re)
        A1 = sp.GpuCsrMat.new(Ab.indptr, Ab.indices, Ab.data, shape=(n, n), ctx=ctx, row_range=(rb, re))
        # ---- (b) on-device generator (this rank's block)
        A2 = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, G, G, G, params=(1.0, 0.5, 0.25), ctx=ctx)
        assert (A2.n_local, A2.row_begin) == (re - rb, rb)
        xg = np.cos(0.37 * np.arange(n)) + 0.25
        xl = xg[rb:re].copy()
        for name, A in (("host", A1), ("stencil", A2)):
            y = np.zeros(re - rb)
            for _ in range(3):  # repeated products exercise the parity double-buffering
                A.mul_vec(xl, y)
            res[f"spmv_{name}"] = y.copy()
            y2 = np.zeros(re - rb)
            res[f"dot_{name}"] = A.mul_vec_dot(xl, y2)
            assert np.array_equal(y2, y)
        # ---- Jacobi-BiCGStab, rhs = A * 1
        ones = np.ones(re - rb)
        rhs = np.zeros(re - rb)
        A2.mul_vec(ones, rhs)
        M = sp.DiagPrecond.from_matrix(A2)
        S = sp.BiCGStab(A2, re - rb).record_history(600)
        x = np.zeros(re - rb)
        it, rr = S.precond_solve(M, rhs, x, 500, 1e-8)
        res["bicg"] = (it, rr, S.history.copy(), x.copy())
        # ---- MINRES on the shifted 7-point system
        A3 = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, G, G, G, params=(0.05,), ctx=ctx)
        rhs3 = np.zeros(re - rb)
        A3.mul_vec(ones, rhs3)
        S3 = sp.MinRes(A3, re - rb).record_history(600)
        x3 = np.zeros(re - rb)
        it3, rr3 = S3.solve(rhs3, x3, 500, 1e-8)
        res["minres"] = (it3, rr3, S3.history.copy(), x3.copy())
        # ---- mv_hint on a partitioned matrix (collective): re-tunes and re-classifies the tiles, results unchanged
        A2.mv_and_dotmv_hint(1500)
        y = np.zeros(re - rb)
        A2.mul_vec(xl, y)
        res["spmv_hinted"] = y.copy()
        S = sp.BiCGStab(A2, re - rb).record_history(600)
        x = np.zeros(re - rb)
        it, rr = S.precond_solve(M, rhs, x, 500, 1e-8)
        res["bicg_hinted"] = (it, rr, S.history.copy(), x.copy())
        # ---- the exact-dot golden case: 27-point 96^3 partitioned (tests/golden/exact_v1.npz)
        G96 = 96
        A96 = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, G96, G96, G96, params=(1.0, 0.5, 0.25), ctx=ctx)
        nl = A96.n_local
        o96 = np.ones(nl)
        r96 = np.zeros(nl)
        A96.mul_vec(o96, r96)
        S96 = sp.BiCGStab(A96, nl).record_history(2000)
        x96 = np.zeros(nl)
        it96, rr96 = S96.precond_solve(sp.DiagPrecond.from_matrix(A96), r96, x96, 2000, 1e-8)
        res["c5_96"] = (it96, rr96, S96.history.copy(), x96.copy())
        # ---- one-directional coupling (block lower bidiagonal: rank r reads rank r-1 only, rank 0 only
        # sends): consecutive products with a changing x and no reduction in between -- the halo put must
        # not overwrite a parity buffer a slower neighbour still gathers from (acknowledged puts, peer.cuh)
        nb = 4096 * world
        lo, hi = nb // world * rank, nb // world * (rank + 1)
        K = nb // world
        rows = np.arange(lo, hi)
        has = rows >= K
        ipb = np.concatenate([[0], np.cumsum(1 + has.astype(np.int64))])
        idxb = np.empty(ipb[-1], np.int32)
        valb = np.empty(ipb[-1])
        pos = ipb[:-1]
        idxb[pos[has]] = rows[has] - K
        valb[pos[has]] = 0.5
        idxb[pos + has] = rows
        valb[pos + has] = 2.0
        AB = sp.GpuCsrMat.new(ipb, idxb, valb, shape=(nb, nb), ctx=ctx, row_range=(lo, hi))
        import torch

        dev = torch.device(f"cuda:{rank}")
        xs = [torch.arange(lo, hi, dtype=torch.float64, device=dev) * (k + 1) + k for k in range(40)]
        ys = [torch.empty(hi - lo, dtype=torch.float64, device=dev) for _ in range(40)]
        torch.cuda.synchronize(dev)
        for k in range(40):
            AB.mul_vec_dev(xs[k].data_ptr(), ys[k].data_ptr())
        ctx.synchronize()
        res["onedir"] = np.stack([y.cpu().numpy() for y in ys])
        res["block"] = (rb, re)
        out.put((rank, res))
        ctx.synchronize()
    except Exception as e:  # pragma: no cover
        import traceback

        out.put((rank, {"error": f"{e}\n{traceback.format_exc()}"}))


@pytest.mark.gpu
@pytest.mark.multigpu
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_partitioned_parity(orc, transport):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    import sprsolve_b200 as sp

    uid = sp.Context.comm_unique_id()
    mpctx = mp.get_context("spawn")
    out = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, world, uid, transport, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert "error" not in got[r], got[r]["error"]
    G = 24
    n = G**3
    A = orc.gen_convdiff27(G)
    xg = np.cos(0.37 * np.arange(n)) + 0.25
    y_ref = orc.spmv(A, xg)
    for name in ("host", "stencil"):
        y = np.concatenate([got[r][f"spmv_{name}"] for r in range(world)])
        assert np.array_equal(y, y_ref), f"partitioned SpMV ({name}) is not bit-exact"
        d_ref = float(np.dot(xg, y_ref))
        for r in range(world):
            assert abs(got[r][f"dot_{name}"] - d_ref) <= 1e-13 * abs(d_ref)
            assert got[r][f"dot_{name}"] == got[0][f"dot_{name}"]  # identical bits on every rank
    # BiCGStab vs the serial oracle: the north-star criteria with the measured re-ordering noise floor
    # of the reference algorithm itself (tests/test_gpu_parity.py::_check_solve)
    from test_gpu_parity import _check_solve, _noise_floor

    rhs = orc.spmv(A, np.ones(n))
    pc = ("diag", A.diagonal())
    o = orc.bicgstab(A, rhs, max_iter=500, tol=1e-8, pc=pc, hist_cap=501)
    assert o.status == orc.OK
    it, rr, hist, _ = got[0]["bicg"]
    for r in range(world):  # every rank carries identical scalars
        assert got[r]["bicg"][0] == it and np.array_equal(got[r]["bicg"][2], hist)
    x = np.concatenate([got[r]["bicg"][3] for r in range(world)])
    floor, it_range = _noise_floor(orc, "bicgstab", A, rhs, pc, 1e-8, 500, o)
    _check_solve(((it, rr, x, hist), o, floor, it_range), 1e-8, A, rhs, xs=np.ones(n))
    # ... and bit for bit the single-GPU solve: SpMV and the vector updates are element-wise
    # identical, the reductions are order-independent (double-double, csrc/reduce.cuh), so the
    # partition of the rows leaves no trace in the iterates.
    G1 = sp.GpuCsrMat.new(A.indptr, A.indices, A.data)
    S1 = sp.BiCGStab(G1, n).record_history(600)
    x1 = np.zeros(n)
    it1, rr1 = S1.precond_solve(sp.DiagPrecond.from_matrix(G1), rhs, x1, 500, 1e-8)
    assert it1 == it and rr1 == rr and np.array_equal(S1.history, hist) and np.array_equal(x1, x)
    # mv_hint on the partitioned matrix changed nothing
    assert np.array_equal(np.concatenate([got[r]["spmv_hinted"] for r in range(world)]), y_ref)
    for r in range(world):
        assert got[r]["bicg_hinted"][0] == it and np.array_equal(got[r]["bicg_hinted"][2], hist)
    assert np.array_equal(np.concatenate([got[r]["bicg_hinted"][3] for r in range(world)]), x)
    # ... and the exact-dot oracle bit for bit, over every iteration, on the partitioned 96^3 system
    gold = np.load(os.path.join(ROOT, "tests", "golden", "exact_v1.npz"))
    it96, rr96, h96, _ = got[0]["c5_96"]
    assert it96 == int(gold["c5_96.exact.iters"]) and rr96 == float(gold["c5_96.exact.resid"])
    assert np.array_equal(h96, gold["c5_96.exact.hist"])
    x96 = np.concatenate([got[r]["c5_96"][3] for r in range(world)])
    import hashlib

    assert hashlib.sha256(x96.tobytes()).hexdigest() == str(gold["c5_96.exact.x_sha256"])
    # one-directional coupling: y_k = 2 x_k + 0.5 x_k[row - K]
    nb = 4096 * world
    K = nb // world
    for k in range(40):
        xk = np.arange(nb, dtype=np.float64) * (k + 1) + k
        yk = 2.0 * xk
        yk[K:] = 0.5 * xk[:-K] + 2.0 * xk[K:]
        assert np.array_equal(np.concatenate([got[r]["onedir"][k] for r in range(world)]), yk), k
    # MINRES vs the serial oracle (well conditioned w.r.t. summation order: strict 1e-10 / +-2 %)
    A3 = orc.gen_lap3d7(G, shift=0.05)
    rhs3 = orc.spmv(A3, np.ones(n))
    o3 = orc.minres(A3, rhs3, max_iter=500, tol=1e-8, hist_cap=501)
    assert o3.status == orc.OK
    it3, rr3, h3, _ = got[0]["minres"]
    x3 = np.concatenate([got[r]["minres"][3] for r in range(world)])
    floor3, range3 = _noise_floor(orc, "minres", A3, rhs3, None, 1e-8, 500, o3)
    _check_solve(((it3, rr3, x3, h3), o3, floor3, range3), 1e-8, A3, rhs3, strict=True, xs=np.ones(n))
    G3 = sp.GpuCsrMat.new(A3.indptr, A3.indices, A3.data)
    S3 = sp.MinRes(G3, n).record_history(600)
    x31 = np.zeros(n)
    it31, rr31 = S3.solve(rhs3, x31, 500, 1e-8)
    assert it31 == it3 and rr31 == rr3 and np.array_equal(S3.history, h3) and np.array_equal(x31, x3)
