"""The reference's own bench sizes on the GPU next to the CPU port (VERDICT r01 "missing #2").

    python tools/bench_small.py [--out profiles/r02_small_systems.jsonl]

Mirrors benches/bicgstab.rs:14-37 (100^2 Dirichlet Laplacian, BiCGStab::solve, tol 1e-16, max 1500 its),
benches/mkl_bicgstab.rs:9-32 (110^2, through MklMat + mv_and_dotmv_hint(1500)) and
benches/mat_vec_mul.rs:15-36 (140^2 SpMV), each with x reset before every solve (the reference's
harness warm-starts after the first sample), plus a size sweep of Jacobi-BiCGStab
(config C1's recipe) to locate the GPU/CPU crossover.  Per case: wall-clock per solve and per
iteration for
  * gpu_fused   : the single-kernel solve (csrc/bicgstab.cu: bicg_fused_kernel), device-resident vectors
  * gpu_multi   : the multi-kernel loop (SPB_FUSED=0)
  * gpu_host    : the reference-facing call on HOST slices (H2D of rhs/x + D2H of x inside)
  * cpu_serial  : oracle port, 1 thread (reference without `parallel`/`mkl`)
  * cpu_rayon4  : oracle port, row-parallel SpMV on 4 threads (the benches' rayon setting, bicgstab.rs:7)
The oracle is used here as the timed CPU arm only (as bench.py's cpu_baseline leg does).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def wall(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    import sprsolve_b200 as sp
    from oracle import oracle as orc

    ctx = sp.default_context()
    dev = torch.device("cuda:0")
    lines = []

    def emit(d):
        lines.append(d)
        print(json.dumps(d), flush=True)

    def bicg_case(name, g, tol, max_iter, jacobi, hint=False):
        A, rhs = orc.gen_dirichlet2d(g)
        G = sp.GpuCsrMat.new(A.indptr.astype(np.int32), A.indices, A.data)
        if hint:
            G.mv_and_dotmv_hint(1500)
        M = sp.DiagPrecond.new(A.diagonal()) if jacobi else None
        S = sp.BiCGStab(G, A.n)
        d_rhs = torch.from_numpy(rhs).to(dev)
        d_x = torch.zeros(A.n, dtype=torch.float64, device=dev)
        out = {"case": name, "grid": g, "n": A.n, "nnz": A.nnz, "tol": tol, "jacobi": jacobi}

        def gpu_dev():
            d_x.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = S.solve_dev(d_rhs.data_ptr(), d_x.data_ptr(), max_iter, tol, precond=M)
            return time.perf_counter() - t0, r

        def timed_dev():
            gpu_dev()
            ts = [gpu_dev() for _ in range(args.reps)]
            return float(np.median([t for t, _ in ts])), ts[0][1]

        os.environ["SPB_FUSED"] = "1"
        t, (it, res) = timed_dev()
        xf = d_x.cpu().numpy().copy()
        out["iters"] = it
        out["gpu_fused"] = {"solve_ms": 1e3 * t, "us_per_iter": 1e6 * t / max(it, 1)}
        os.environ["SPB_FUSED"] = "0"
        t, (it2, res2) = timed_dev()
        out["gpu_multi"] = {"solve_ms": 1e3 * t, "us_per_iter": 1e6 * t / max(it2, 1)}
        out["fused_equals_multi"] = bool(it2 == it and res2 == res and np.array_equal(xf, d_x.cpu().numpy()))
        del os.environ["SPB_FUSED"]
        x = np.zeros(A.n)

        def gpu_host():
            x[:] = 0.0
            return S.precond_solve(M, rhs, x, max_iter, tol) if M is not None else S.solve(rhs, x, max_iter, tol)

        t, _ = wall(gpu_host, args.reps)
        out["gpu_host"] = {"solve_ms": 1e3 * t, "us_per_iter": 1e6 * t / max(it, 1)}
        pc = ("diag", A.diagonal()) if jacobi else None
        for tag, mode, thr in (("cpu_serial", 0, 1), ("cpu_rayon4", 1, 4)):
            orc.set_mode(mode)
            orc.set_threads(thr)
            t, o = wall(lambda: orc.bicgstab(A, rhs, max_iter=max_iter, tol=tol, pc=pc, hist_cap=1), max(2, args.reps // 2))
            out[tag] = {"solve_ms": 1e3 * t, "us_per_iter": 1e6 * t / max(o.iters, 1), "iters": o.iters, "threads": thr}
        orc.set_mode(0)
        orc.set_threads(orc.max_threads())
        out["gpu_fused_vs_cpu_rayon4"] = out["cpu_rayon4"]["us_per_iter"] / out["gpu_fused"]["us_per_iter"]
        emit(out)

    # the reference's benches
    bicg_case("benches/bicgstab.rs (100^2, solve, tol 1e-16)", 100, 1e-16, 1500, False)
    bicg_case("benches/mkl_bicgstab.rs (110^2, MklMat + mv_and_dotmv_hint)", 110, 1e-16, 1500, False, hint=True)
    # Jacobi-BiCGStab sweep (C1's recipe): where does the GPU overtake the CPU port?
    for g in (32, 64, 100, 140, 256, 512):
        bicg_case(f"C1 recipe {g}^2 (Jacobi, rtol 1e-8)", g, 1e-8, 10000, True)

    # benches/mat_vec_mul.rs: 140^2 SpMV
    for g in (140, 512):
        A, rhs = orc.gen_dirichlet2d(g)
        G = sp.GpuCsrMat.new(A.indptr.astype(np.int32), A.indices, A.data)
        y = np.zeros(A.n)
        t_host, _ = wall(lambda: G.mul_vec(rhs, y), 20)
        d_in = torch.from_numpy(rhs).to(dev)
        d_out = torch.empty_like(d_in)
        torch.cuda.synchronize()
        reps = 2000
        for _ in range(10):
            G.mul_vec_dev(d_in.data_ptr(), d_out.data_ptr())
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            G.mul_vec_dev(d_in.data_ptr(), d_out.data_ptr())
        ctx.synchronize()
        t_dev = (time.perf_counter() - t0) / reps
        out = {"case": f"benches/mat_vec_mul.rs ({g}^2 SpMV)", "grid": g, "n": A.n, "nnz": A.nnz,
               "gpu_host_us": 1e6 * t_host, "gpu_dev_us": 1e6 * t_dev}
        for tag, par, thr in (("cpu_serial_us", False, 1), ("cpu_rayon4_us", True, 4)):
            orc.set_threads(thr)
            t, _ = wall(lambda: orc.spmv(A, rhs, parallel=par), 50)
            out[tag] = 1e6 * t
        orc.set_threads(orc.max_threads())
        emit(out)

    if args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
