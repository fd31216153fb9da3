#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not the bench.py line (C1, C2, C3, C4) on one B200:

    python tools/bench_configs.py [--configs c1,c2,c3,c4] [--no-cpu] > profiles/rNN_configs.jsonl

Per config one JSON line: full-solve iterations / seconds / iterations per second on the GPU
(device-resident vectors, CUDA events), the per-kernel-family CUDA-event split of one profiled
solve, the algorithmic-bytes model of an iteration and the GB/s it implies, a size-independent
correctness property (known solution), and the oracle port timed on the host cores on the same
matrix (or a bounded sample of it).  Uses the public Python mirror of the reference interface.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def spmv_bytes(n, nnz, val_bytes, ip_bytes=4):
    return nnz * (val_bytes + 4) + (n + 1) * ip_bytes + 2 * n * val_bytes


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def family_split(ctx, fn):
    ctx.profile_reset()
    ctx.profile(True)
    fn()
    torch.cuda.synchronize()
    names = ["spmv", "vector", "scalar", "precond", "pack"]
    out = {}
    for i, nm in enumerate(names):
        n, ms = ctx.profile_read(i)
        out[nm] = {"launches": n, "ms": ms}
    ctx.profile(False)
    ctx.profile_reset()
    return out


def solve_case(ctx, dev, name, A, M, solver_cls, rhs, xs, tol, max_iter, iter_bytes, reps=3):
    n = A.n_local
    S = solver_cls(A, n)
    x = torch.zeros_like(rhs)

    def run():
        x.zero_()
        try:
            return S.solve_dev(rhs.data_ptr(), x.data_ptr(), max_iter, tol, precond=M)
        except sp.SolverError as e:
            return (max_iter, float("nan"), type(e).__name__)

    run()  # warm-up
    best = None
    for _ in range(reps):
        out, ms = timed(run)
        if best is None or ms < best[1]:
            best = (out, ms)
    out, ms = best
    its = out[0]
    its_done = its + 1 if solver_cls in (sp.MinRes, sp.CSMinRes) else its  # MINRES index is 0-based
    err = float((x - xs).abs().max().item()) if xs is not None else None
    launches0 = ctx.launch_count
    split = family_split(ctx, run)
    launches = ctx.launch_count - launches0
    ips = its_done / (ms / 1e3)
    line = {
        "config": name, "rows": n, "nnz": A.nnz, "iterations": its, "rel_residual": out[1], "status": "converged" if len(out) == 2 else out[2],
        "solve_ms": ms, "iters_per_s": ips, "us_per_iter": 1e3 * ms / max(its_done, 1), "max_abs_err_vs_known_solution": err,
        "iter_bytes_model": iter_bytes, "iter_gbs": iter_bytes * ips / 1e9, "iter_frac_of_peak": iter_bytes * ips / 1e9 / PEAK,
        "family_split_one_solve": split, "launches_one_solve": launches,
    }
    S.destroy()
    return line


def cpu_solve(kind, A, rhs, pc, tol, max_iter, threads):
    from oracle import oracle as orc

    orc.build()
    orc.set_threads(threads)
    orc.set_mode(2)
    try:
        t0 = time.perf_counter()
        if kind == "csminres":
            o = orc.csminres(A, rhs, max_iter=max_iter, tol=tol)
        else:
            o = getattr(orc, kind)(A, rhs, max_iter=max_iter, tol=tol, pc=pc)
        dt = time.perf_counter() - t0
    finally:
        orc.set_mode(0)
    done = o.iters + (1 if kind != "bicgstab" else 0) if o.status == orc.OK else max_iter
    return {"iterations": int(o.iters), "status": int(o.status), "seconds": dt, "iters_per_s": done / dt, "cores": threads, "kind": "port"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--c1-grid", type=int, default=512)
    ap.add_argument("--c3-grid", type=int, default=128)
    ap.add_argument("--c4-grid", type=int, default=200)
    args = ap.parse_args()
    todo = args.configs.split(",")
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = sp.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    threads = os.cpu_count() or 1
    f64, c128 = torch.float64, torch.complex128

    if "c1" in todo:
        g = args.c1_grid
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g, g, 1, ctx=ctx)
        n = A.n_local
        M = sp.DiagPrecond.from_matrix(A)
        # rhs = i + j on the border, 0 inside (src/main.rs:9-11, 90-103)
        ii = torch.arange(g, device=dev, dtype=f64)
        R = ii[:, None] + ii[None, :]
        mask = torch.zeros(g, g, dtype=torch.bool, device=dev)
        mask[0, :] = mask[-1, :] = mask[:, 0] = mask[:, -1] = True
        rhs = torch.where(mask, R, torch.zeros_like(R)).reshape(-1).contiguous()
        xs = R.reshape(-1).contiguous()  # harmonic: u = i + j solves the discrete Laplace problem exactly
        ib = 2 * spmv_bytes(n, A.nnz, 8) + 21 * n * 8
        line = solve_case(ctx, dev, f"C1 Jacobi-BiCGStab f64 2-D 5-pt Dirichlet {g}^2 rtol 1e-8", A, M, sp.BiCGStab, rhs, xs, 1e-8, 20000, ib)
        line["note"] = "21 MB working set: L2-resident and launch/latency-bound, not graded against the HBM roofline"
        if not args.no_cpu:
            from oracle import oracle as orc

            Ac, rc = orc.gen_dirichlet2d(g)
            line["cpu_baseline"] = cpu_solve("bicgstab", Ac, rc, ("diag", Ac.diagonal()), 1e-8, 20000, threads)
            line["cpu_baseline_4threads"] = cpu_solve("bicgstab", Ac, rc, ("diag", Ac.diagonal()), 1e-8, 20000, min(4, threads))
        print(json.dumps(line), flush=True)
        del A, M

    if "c2" in todo:
        g = 256
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.0,), ctx=ctx)
        n = A.n_local
        k = torch.arange(n, device=dev)
        x = 1.0 + (k % 17).double() / 17.0
        y = torch.empty_like(x)
        for _ in range(5):
            A.mul_vec_dev(x.data_ptr(), y.data_ptr())
        reps = 50
        _, ms = timed(lambda: [A.mul_vec_dev(x.data_ptr(), y.data_ptr()) for _ in range(reps)])
        b = spmv_bytes(n, A.nnz, 8)
        # size-independent property: row sums of the Laplacian -> A * 1 = (6 - #neighbours)
        ones = torch.ones_like(x)
        A.mul_vec_dev(ones.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        interior_zero = bool((y.reshape(g, g, g)[1:-1, 1:-1, 1:-1] == 0).all().item())
        line = {"config": "C2 CSR SpMV f64 3-D 7-pt 256^3", "rows": n, "nnz": A.nnz, "ms": ms / reps, "bytes": b,
                "gbs": b / (ms / reps * 1e-3) / 1e9, "frac_of_peak": b / (ms / reps * 1e-3) / 1e9 / PEAK, "A_times_ones_interior_is_zero": interior_zero}
        if not args.no_cpu:
            from oracle import oracle as orc

            Ac = orc.gen_lap3d7(g)
            xc = x.cpu().numpy()
            orc.set_threads(threads)
            orc.spmv(Ac, xc, parallel=True)
            t0 = time.perf_counter()
            for _ in range(5):
                yc = orc.spmv(Ac, xc, parallel=True)
            dt = (time.perf_counter() - t0) / 5
            A.mul_vec_dev(x.data_ptr(), y.data_ptr())
            torch.cuda.synchronize()
            line["bit_exact_vs_oracle_full_size"] = bool(np.array_equal(yc, y.cpu().numpy()))
            line["cpu_baseline"] = {"ms": dt * 1e3, "gbs": b / dt / 1e9, "cores": threads, "kind": "port",
                                    "note": "oracle SpMV, OpenMP row chunks >= 128 (src/mat.rs:85-107); same byte formula (i32 columns)"}
        print(json.dumps(line), flush=True)
        del A

    if "c3" in todo:
        g = args.c3_grid
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.05,), ctx=ctx)
        n = A.n_local
        t0 = time.perf_counter()
        M = sp.GaussSeidelPrecond(A, symmetric=True)
        t_setup = time.perf_counter() - t0
        xs = torch.ones(n, dtype=f64, device=dev)
        rhs = torch.empty_like(xs)
        A.mul_vec_dev(xs.data_ptr(), rhs.data_ptr())
        sgs_bytes = 2 * (A.nnz * 12 + 4 * n * 8)
        ib = spmv_bytes(n, A.nnz, 8) + 13 * n * 8 + sgs_bytes
        line = solve_case(ctx, dev, f"C3 SGS-MINRES f64 shifted 3-D 7-pt {g}^3 sigma=0.05 rtol 1e-8", A, M, sp.MinRes, rhs, xs, 1e-8, 5000, ib)
        line["gs_levels_fwd_bwd"] = M.levels()
        line["gs_setup_seconds"] = t_setup
        line2 = solve_case(ctx, dev, f"C3 (unpreconditioned MINRES, same matrix)", A, None, sp.MinRes, rhs, xs, 1e-8, 5000, spmv_bytes(n, A.nnz, 8) + 13 * n * 8)
        if not args.no_cpu:
            from oracle import oracle as orc

            Ac = orc.gen_lap3d7(g, shift=0.05)
            rc = orc.spmv(Ac, np.ones(Ac.n))
            line["cpu_baseline"] = cpu_solve("minres", Ac, rc, ("gs_sym",), 1e-8, 5000, threads)
        print(json.dumps(line), flush=True)
        print(json.dumps(line2), flush=True)
        del A, M

    if "c4" in todo:
        g = args.c4_grid
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.5, 0.5), dtype=np.complex128, ctx=ctx)
        n = A.n_local
        xs = torch.full((n,), 1 + 1j, dtype=c128, device=dev)
        rhs = torch.empty_like(xs)
        A.mul_vec_dev(xs.data_ptr(), rhs.data_ptr())
        ib = spmv_bytes(n, A.nnz, 16) + 13 * n * 16
        line = solve_case(ctx, dev, f"C4 CSMinRes complex128 Helmholtz 7-pt {g}^3 k2=0.5 gamma=0.5 rtol 1e-8", A, None, sp.CSMinRes, rhs, xs, 1e-8, 5000, ib)
        # standalone complex SpMV
        y = torch.empty_like(xs)
        for _ in range(3):
            A.mul_vec_dev(xs.data_ptr(), y.data_ptr())
        _, ms = timed(lambda: [A.mul_vec_dev(xs.data_ptr(), y.data_ptr()) for _ in range(30)])
        b = spmv_bytes(n, A.nnz, 16)
        line["spmv_z"] = {"ms": ms / 30, "bytes": b, "gbs": b / (ms / 30 * 1e-3) / 1e9, "frac_of_peak": b / (ms / 30 * 1e-3) / 1e9 / PEAK}
        if not args.no_cpu:
            from oracle import oracle as orc

            gs = min(g, 128)  # bounded sample
            Ac = orc.gen_lap3d7(gs, shift=0.5 + 0.5j, dtype=np.complex128)
            rc = orc.spmv(Ac, np.full(Ac.n, 1 + 1j))
            cb = cpu_solve("csminres", Ac, rc, None, 1e-8, 5000, threads)
            cb["sample"] = f"{gs}^3 of the same family; iterations/s scaled by rows to {g}^3"
            cb["iters_per_s_scaled"] = cb["iters_per_s"] * (gs / g) ** 3
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
