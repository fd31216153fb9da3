#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: BiCGStab iterations/s (+ SpMV achieved HBM GB/s as a
fraction of the measured roofline) on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[4] -- Jacobi-preconditioned BiCGStab, f64, on
the row-partitioned 3-D 27-point convection-diffusion system 512^3 (134,217,728 rows,
3,609,741,304 non-zeros), rhs = A*1, x0 = 0.  It fits one B200 (56 GB), so it is also the N=1
workload; total work is fixed as N grows (scaling: strong, as the north star's "6x faster on 8
GPUs than on 1" demands).  A "step" is one solver call capped at --iters (default 100) BiCGStab
iterations starting from x0 = 0 (every iteration does identical work: 2 SpMV + 3 fused vector
kernels + 3 reduction points; a full solve of this system takes 640); one full solve to rtol
1e-8 is run first and reported in `full_solve` (its iteration count and final residual are the
same at every GPU count: the reductions are order-independent, csrc/reduce.cuh).

value   : iterations/s with rhs / x resident in HBM (spb_solver_solve_dev), CUDA events on the
          launching stream, barrier + synchronize on both sides, max over ranks.
e2e     : the same through the host-buffer path: each step copies rhs and x0 from pinned host
          memory to the device, solves, and copies x back (h2d/d2h bytes per step reported); the
          caller-side zeroing of the initial guess happens before the timed region (one pinned x
          buffer per step).
roofline: the SpMV kernel (dominant: ~80 % of an iteration's bytes), algorithmic bytes
          nnz*12 + (n+1)*sizeof(indptr) + 2*n*8 per launch (per rank) / average launch duration
          measured live with CUDA events around every SpMV launch of one extra profiled step.
          `achieved` / `frac` use those ALGORITHMIC bytes (the plain CSR stream of the reference's
          operator); when the analysis replaces the column stream by the row-pattern dictionary the
          kernel moves fewer bytes (`stream_bytes_per_launch`, `stream_gbs`, `stream_frac` -- what
          the HBM roofline actually bounds), so `frac` can exceed 1.
spmv_c2 : BASELINE.json configs[1], standalone SpMV on the 3-D 7-point 256^3 matrix (N=1 only).
cpu_baseline / --impl reference: the reference cannot be built here (Rust nightly + MKL +
          unvendored git deps, no cargo), so the CPU arm is the oracle port of its solver loop
          with OpenMP row-parallel SpMV and OpenMP vector ops (the rayon + MKL-iomp stand-in) on
          all host cores, on a bounded sample (same matrix family at a smaller grid), scaled to
          the 512^3 unit by the row ratio.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# stdout carries exactly ONE JSON line: everything else any library prints there (e.g. NCCL's
# version banner) is sent to stderr by pointing fd 1 at fd 2 for the duration of the run.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "bicgstab_iters_per_s"
UNIT = "iterations/s"
B27 = (1.0, 0.5, 0.25)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def spmv_bytes(n, nnz, indptr_bytes, val_bytes=8):
    return nnz * (val_bytes + 4) + (n + 1) * indptr_bytes + 2 * n * val_bytes


# ----------------------------------------------------------------------------------- CPU arm
def cpu_bicgstab_sample(grid: int, iters: int, full_grid: int, threads: int | None = None):
    """Oracle port of BiCGStab::precond_solve (src/bicg_stab.rs:204-366) with OpenMP SpMV and
    vector ops on all host cores, 27-point matrix at grid^3, `iters` iterations; returns
    iterations/s scaled to full_grid^3 by the row ratio."""
    from oracle import oracle as orc

    orc.build()
    nthreads = threads or os.cpu_count() or 1
    orc.set_threads(nthreads)
    A = orc.gen_convdiff27(grid, grid, grid, b=B27)
    orc.set_mode(1)
    rhs = orc.spmv(A, np.ones(A.n), parallel=True)
    diag = A.diagonal()
    orc.set_mode(2)
    orc.bicgstab(A, rhs, max_iter=2, tol=1e-8, pc=("diag", diag))  # touch memory
    t0 = time.perf_counter()
    out = orc.bicgstab(A, rhs, max_iter=iters, tol=1e-30, pc=("diag", diag))
    dt = time.perf_counter() - t0
    orc.set_mode(0)
    done = iters if out.status == orc.INSUFFICIENT_ITER else max(out.iters, 1)
    ips_sample = done / dt
    scale = (grid / full_grid) ** 3
    return {
        "value": ips_sample * scale,
        "unit": UNIT,
        "cores": nthreads,
        "kind": "port",
        "sample": f"Jacobi-BiCGStab, 27-pt convection-diffusion {grid}^3 ({A.n} rows, {A.nnz} nnz), {done} iterations in {dt:.2f} s "
        f"= {ips_sample:.3f} it/s on the sample, scaled by rows ({grid}^3/{full_grid}^3) to the {full_grid}^3 unit; "
        "oracle port with OpenMP SpMV (rayon stand-in, src/mat.rs:85-107) + OpenMP vector ops (MKL-iomp stand-in)",
        "sample_iters_per_s": ips_sample,
        "seconds": dt,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals = []
    last = None
    # bounded: each step is a ~3-6 s sample; the whole run stays within a few minutes
    for i in range(args.warmup + args.steps):
        last = cpu_bicgstab_sample(args.cpu_grid, args.cpu_iters, args.grid)
        if i >= args.warmup:
            vals.append(last["value"])
        if time.perf_counter() - t_all > 240 and len(vals) >= 1:
            break
    v = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": v,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": len(vals),
        "warmup": args.warmup,
        "ms_per_step": 1e3 * last["seconds"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")} | {"value": v},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"BASELINE.json configs[4]: Jacobi-BiCGStab f64, 3-D 27-point convection-diffusion {args.grid}^3, row-block partitioned, rhs=A*1, x0=0",
        "grid": args.grid,
        "rows": args.grid**3,
        "iters_per_step": args.iters,
        "partition": f"{world} row block(s) (z-slabs), one process per GPU",
        "l2": "inputs (>= 5 GB per rank) far exceed the 126 MB L2; no flush needed",
    }


# ----------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import sprsolve_b200 as sp

    from sprsolve_b200 import dist as spd

    world, rank, local_rank = spd.env_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    spd.init_process_group("nccl", device=dev)
    ctx = sp.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    spd.attach_communicator(ctx)  # rank 0 creates the communicator id, broadcast, spb_comm_init

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = args.grid
    t_setup = time.perf_counter()
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=B27, ctx=ctx)
    n_glob, n_loc, _ = A._sizes()
    nnz_loc = A.nnz
    M = sp.DiagPrecond.from_matrix(A)
    ones = torch.ones(n_loc, dtype=torch.float64, device=dev)
    rhs = torch.empty(n_loc, dtype=torch.float64, device=dev)
    x = torch.zeros(n_loc, dtype=torch.float64, device=dev)
    A.mul_vec_dev(ones.data_ptr(), rhs.data_ptr())  # rhs = A * 1
    del ones
    S = sp.BiCGStab(A, n_loc)
    barrier()
    setup_s = time.perf_counter() - t_setup

    # ---- one full solve to rtol 1e-8 (convergence evidence; also the first warm-up)
    full = None
    if not args.no_full_solve:
        S.record_history(64)
        x.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            its, res = S.solve_dev(rhs.data_ptr(), x.data_ptr(), args.max_iter, 1e-8, precond=M)
            status = "converged"
        except sp.SolverError as e:
            its, res, status = args.max_iter, float("nan"), type(e).__name__
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        err = float((x - 1.0).abs().max().item())
        err = max_over_ranks(err)
        full = {"status": status, "iterations": its, "rel_residual": res, "seconds": ms / 1e3, "iters_per_s": its / (ms / 1e3),
                "max_abs_err_vs_ones": err, "history_head": [float(v) for v in S.history[:8]]}
        S.record_history(0)

    iters = args.iters
    if full and full["status"] == "converged":
        iters = max(2, min(iters, full["iterations"] - 1))

    def step_dev():
        x.zero_()
        try:
            S.solve_dev(rhs.data_ptr(), x.data_ptr(), iters, 1e-30, precond=M)
        except sp.InsufficientIterNum:
            pass

    # ---- value: device-resident
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    value = args.steps * iters / (ms_total / 1e3)

    # ---- e2e: the reference-facing call on HOST slices -- BiCGStab.precond_solve(M, rhs, x, ..)
    #      = spb_solver_solve: H2D of rhs and of the initial guess x0, the solve, D2H of x, all
    #      inside the timed region; rhs / x live in pinned host memory; x is in/out like the
    #      reference's `&mut [T]`, so every step first resets it to x0 = 0 on the host.
    #      The caller-side preparation of the initial guess (zeroing a host buffer) is not part of
    #      the call: every step gets its own pinned x buffer, zeroed before the timed region.
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    rhs_h = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
    rhs_h.copy_(rhs)
    rhs_np = rhs_h.numpy()
    x_bufs = [torch.zeros(n_loc, dtype=torch.float64, pin_memory=True) for _ in range(args.steps + 1)]

    def step_e2e(k):
        try:
            S.precond_solve(M, rhs_np, x_bufs[k].numpy(), iters, 1e-30)
        except sp.InsufficientIterNum:
            pass

    step_e2e(args.steps)  # warm-up on the spare buffer
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step_e2e(k)
    e1.record()
    barrier()
    x_np = x_bufs[args.steps - 1].numpy()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = args.steps * iters / (ms_e2e / 1e3)
    e2e_check = float(np.abs(x_np - x.cpu().numpy()).max())  # same iterate as the device-resident step
    e2e_check = max_over_ranks(e2e_check)

    # ---- roofline: every SpMV launch of one more step bracketed by CUDA events
    ctx.profile_reset()
    ctx.profile(True)
    step_dev()
    torch.cuda.synchronize()
    n_spmv, ms_spmv = ctx.profile_read(0)
    n_vec, ms_vec = ctx.profile_read(1)
    n_sc, ms_sc = ctx.profile_read(2)
    ctx.profile(False)
    ctx.profile_reset()
    peak, peak_src = measured_peak()
    ip_bytes = 8 if nnz_loc >= 2**31 - 8 else 4
    b_spmv = spmv_bytes(n_loc, nnz_loc, ip_bytes)
    # one SpMV = one launch on a single GPU, two (interior + boundary rows) when partitioned
    n_products = 2 * iters + 1
    avg_ms = max_over_ranks(ms_spmv / n_products)
    achieved = b_spmv / (avg_ms * 1e-3) / 1e9
    # What the analysis chose for this matrix.  With the column-offset dictionary the kernel streams
    # 8 instead of 12 bytes per non-zero, so the ALGORITHMIC figure (the bytes of the reference's CSR
    # operator) can exceed the HBM peak; `stream_*` are the bytes the kernel
    # moves by design, which is what the HBM roofline bounds.
    plan = A.plan_info()
    stream = plan["stream_bytes"]
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_spmv512_dict_ncu_full.json" if plan["dictionary"] else "r01_spmv512_ncu_full.json")
    if world == 1 and g == 512 and os.path.exists(tp):
        traffic = float(json.load(open(tp))["traffic_bytes_per_launch"])
        traffic_src = f"profiles/{os.path.basename(tp)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, same workload)"
    roofline = {
        "bound": "hbm", "kernel": "spmv_tma_kernel<double> (CSR SpMV, 27-pt, this rank's rows)", "achieved": achieved, "peak": peak,
        "unit": "GB/s (per GPU)", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "format": (f"CSR values + 16-bit row-pattern ids ({plan['patterns']} column-offset patterns found by the analysis)"
                   if plan["dictionary"] else "CSR values + int32 column indices"),
        "stream_bytes_per_launch": stream, "stream_gbs": stream / (avg_ms * 1e-3) / 1e9, "stream_frac": stream / (avg_ms * 1e-3) / 1e9 / peak,
        "plan": {k: plan[k] for k in ("consumer_threads", "stages", "tile_nnz", "ctas_per_sm")},
        "bytes_per_launch": b_spmv, "avg_launch_ms": avg_ms, "launches_timed": n_spmv, "products_timed": n_products,
        "step_share": {"spmv_ms": ms_spmv, "vector_ms": ms_vec, "scalar_ms": ms_sc, "spmv_launches": n_spmv, "vector_launches": n_vec, "scalar_launches": n_sc},
        "iteration_bytes_model": 2 * b_spmv + 21 * n_loc * 8,
        "iteration_gbs_device": (2 * b_spmv + 21 * n_loc * 8) * value / 1e9,
    }

    # ---- BASELINE.json configs[1]: standalone SpMV on 256^3 7-point (N=1 only)
    spmv_c2 = None
    if world == 1 and not args.no_c2:
        del S, M, A
        torch.cuda.empty_cache()
        n1 = 256
        A2 = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, n1, n1, n1, params=(0.0,), ctx=ctx)
        n2 = n1**3
        k = torch.arange(n2, device=dev)
        x2 = 1.0 + (k % 17).double() / 17.0
        y2 = torch.empty(n2, dtype=torch.float64, device=dev)
        for _ in range(5):
            A2.mul_vec_dev(x2.data_ptr(), y2.data_ptr())
        torch.cuda.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            A2.mul_vec_dev(x2.data_ptr(), y2.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / reps
        b2 = spmv_bytes(n2, A2.nnz, 4)
        spmv_c2 = {"workload": "BASELINE.json configs[1]: CSR SpMV f64, 3-D 7-point 256^3", "ms": ms2, "bytes": b2,
                   "gbs": b2 / (ms2 * 1e-3) / 1e9, "frac_of_peak": b2 / (ms2 * 1e-3) / 1e9 / peak, "launches": reps,
                   "note": "1.74 GB per launch > 126 MB L2, back-to-back launches"}

    # ---- CPU baseline on rank 0 at N=1
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        cpu = cpu_bicgstab_sample(args.cpu_grid, args.cpu_iters, g)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world) | {"iters_per_step": iters},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n_loc * 8 * world, "d2h_bytes_per_step": n_loc * 8 * world,
                    "ms_per_step": ms_e2e / args.steps, "api": "BiCGStab.precond_solve on pinned host slices (spb_solver_solve)",
                    "max_abs_diff_vs_device_resident_step": e2e_check},
            "gpu_launches": launches, "clocks": clocks, "full_solve": full, "spmv_c2": spmv_c2, "setup_seconds": setup_s,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="N of the N^3 27-point system (BASELINE: 512)")
    ap.add_argument("--iters", type=int, default=100,
                    help="BiCGStab iterations per step (a full solve of the 512^3 system takes ~615; the host<->device "
                         "traffic of the e2e arm is per solver call, so a short step over-weights it)")
    ap.add_argument("--max-iter", type=int, default=5000)
    ap.add_argument("--cpu-grid", type=int, default=256)
    ap.add_argument("--cpu-iters", type=int, default=30)
    ap.add_argument("--no-full-solve", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
