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
roofline: the SpMV kernel (dominant: ~80 % of an iteration's bytes).  `achieved` = the ALGORITHMIC
          bytes of the operand format the kernel consumes / the average launch
          duration measured live with CUDA events around every SpMV launch of one extra profiled
          step.  Plain CSR: nnz*12 + (n+1)*sizeof(indptr) + 2*n*8.  When the
          analysis finds the column-offset dictionary (27-point matrix: 27 row patterns) the column
          stream and the row-pointer stream do not exist: nnz*8 + (n+1)*4 (row words: pattern id + low 16
          bits of the row pointer) + 2*n*8 -- these
          are the bytes the HBM roofline bounds, so `frac` = achieved/peak stays a physical fraction.
          `csr_equivalent_*` quote the same launch in the reference operator's CSR bytes (can exceed
          the peak: a speed-up in format, not in bandwidth).  `alone` = the same kernel timed outside the
          solve (burst clocks, like the measured copy peak; the long solve runs under the power cap).
spmv_c2 : BASELINE.json configs[1], standalone SpMV on the 3-D 7-point 256^3 matrix (N=1 only); x / y
          rotate over 4 buffer pairs so no launch finds its x (134 MB, L2 evict_last) in the 126 MB L2.
configs : (N=1 only, a few seconds) the other BASELINE configs at FULL size through the same ABI: C1
          512^2 Jacobi-BiCGStab (single-kernel solve), C3 128^3 SGS-MINRES, C4 200^3 CSMINRES --
          iterations/s, and whether iteration count and the whole residual history equal the committed
          exact-dot goldens (tests/golden/exact_v1.npz) bit for bit.
parity_check: at every N, before timing: the 27-point 96^3 system, partitioned like the workload,
          solved to 1e-8 and compared with the exact-dot golden history bit for bit.
cpu_baseline / --impl reference: the reference cannot be built here (Rust nightly + MKL +
          unvendored git deps, no cargo), so the CPU arm is the oracle port of its solver loop
          with OpenMP row-parallel SpMV and OpenMP vector ops (the rayon + MKL-iomp stand-in) on
          all host cores, on a bounded sample (same matrix family at a smaller grid), scaled to
          the 512^3 unit by the row ratio; `extrapolation_check` times two grids (192^3, 256^3) in the
          same run: per-row cost must agree for the scaling to be valid.
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


def bind_to_gpu_numa(local_rank: int):
    """Multi-process runs: keep this rank (and its pinned host buffers) on the CPUs next to its GPU.
    8 unbound processes streaming 400 MB per step through one socket is what held the e2e arm at
    24.5 ms of copies per step at N=8 in round 1.  Returns the cpulist string or None."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus and len(cpus) < (os.cpu_count() or 0):
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def exact_gold():
    p = os.path.join(ROOT, "tests", "golden", "exact_v1.npz")
    return np.load(p) if os.path.exists(p) else None


def gold_match(gold, name, its, res, hist):
    """Bit-for-bit agreement with the committed exact-dot history (a fixture, not the oracle)."""
    if gold is None or f"{name}.exact.hist" not in gold:
        return None
    gh = gold[f"{name}.exact.hist"]
    return {"case": name, "iterations": int(its), "golden_iterations": int(gold[f"{name}.exact.iters"]),
            "residual_equal": bool(res == float(gold[f"{name}.exact.resid"])),
            "history_bit_identical": bool(len(hist) == len(gh) and np.array_equal(np.asarray(hist), gh)), "history_len": int(len(gh))}


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
    # validity of the row-ratio scaling: a second grid in the same run (per-row cost must agree)
    small = cpu_bicgstab_sample(192, args.cpu_iters, args.grid)
    extrap = {"grid_a": 192, "value_a": small["value"], "grid_b": args.cpu_grid, "value_b": v, "ratio_a_over_b": small["value"] / v,
              "note": "both scaled to the 512^3 unit by rows; a ratio near 1 means the per-row cost does not depend on the sample size "
                      "(both samples are far larger than the CPU caches)"}
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
        "extrapolation_check": extrap,
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
    cpu_bind = bind_to_gpu_numa(local_rank) if world > 1 else None
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

    gold = exact_gold()

    # ---- parity before timing (every N): 27-point 96^3, partitioned like the workload, vs the exact-dot golden
    parity = None
    if not args.no_parity:
        Ap = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, 96, 96, 96, params=B27, ctx=ctx)
        npl = Ap.n_local
        op_ = torch.ones(npl, dtype=torch.float64, device=dev)
        rp_ = torch.empty(npl, dtype=torch.float64, device=dev)
        xp_ = torch.zeros(npl, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        Ap.mul_vec_dev(op_.data_ptr(), rp_.data_ptr())
        Sp = sp.BiCGStab(Ap, npl).record_history(2048)
        itp, resp = Sp.solve_dev(rp_.data_ptr(), xp_.data_ptr(), 2000, 1e-8, precond=sp.DiagPrecond.from_matrix(Ap))
        parity = gold_match(gold, "c5_96", itp, resp, Sp.history)
        if parity is not None:
            parity["ranks"] = world
            parity["max_abs_err_vs_ones"] = max_over_ranks(float((xp_ - 1.0).abs().max().item()))
        del Sp, Ap, op_, rp_, xp_
        torch.cuda.empty_cache()

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
    # What the analysis chose for this matrix.  `achieved` counts the algorithmic bytes of the operand
    # format the kernel consumes: with the column-offset dictionary there is no column
    # stream (8 bytes per non-zero + 4 per row instead of 12 + 4..8 per row), which is what the HBM roofline bounds.
    # The same launch in the reference operator's CSR bytes is `csr_equivalent_*`.
    plan = A.plan_info()
    stream = plan["stream_bytes"]
    achieved = stream / (avg_ms * 1e-3) / 1e9
    csr_eq = b_spmv / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    if world == 1 and g == 512:
        stem = "spmv512_dict_ncu_full.json" if plan["dictionary"] else "spmv512_ncu_full.json"
        for rnd in ("r02", "r01"):
            tp = os.path.join(ROOT, "profiles", f"{rnd}_{stem}")
            if os.path.exists(tp):
                meta = json.load(open(tp))
                traffic = float(meta["traffic_bytes_per_launch"])
                traffic_src = (f"profiles/{os.path.basename(tp)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, same workload"
                               + (f", captured at commit {meta['commit']}" if "commit" in meta else "") + ")")
                break
    roofline = {
        "bound": "hbm", "kernel": "spmv_tma_kernel<double> (CSR SpMV, 27-pt, this rank's rows)", "achieved": achieved, "peak": peak,
        "unit": "GB/s (per GPU)", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "frac_basis": "algorithmic bytes of the operand format the kernel consumes (bytes_per_launch)",
        "format": (f"CSR values + one 32-bit row word (16-bit pattern id + low 16 bits of the row pointer; {plan['patterns']} column-offset patterns found by the analysis)"
                   if plan["dictionary"] else "CSR values + int32 column indices"),
        "csr_equivalent_bytes_per_launch": b_spmv, "csr_equivalent_gbs": csr_eq, "csr_equivalent_frac": csr_eq / peak,
        "plan": {k: plan[k] for k in ("consumer_threads", "stages", "tile_nnz", "ctas_per_sm")},
        "bytes_per_launch": stream, "avg_launch_ms": avg_ms, "launches_timed": n_spmv, "products_timed": n_products,
        # the launches inside the solve are the epilogue variants: they also read the dot-product operand (r0 / r, one more
        # n-vector) that `bytes_per_launch` -- the bare product -- does not count; `frac` stays the conservative figure
        "epilogue_operand_bytes_per_launch": n_loc * 8,
        "frac_incl_epilogue_operand": (stream + n_loc * 8) / (avg_ms * 1e-3) / 1e9 / peak,
        "step_share": {"spmv_ms": ms_spmv, "vector_ms": ms_vec, "scalar_ms": ms_sc, "spmv_launches": n_spmv, "vector_launches": n_vec, "scalar_launches": n_sc},
        "iteration_bytes_model": 2 * stream + 21 * n_loc * 8,
        "iteration_gbs_device": (2 * stream + 21 * n_loc * 8) * value / 1e9,
        "iteration_frac_of_peak": (2 * stream + 21 * n_loc * 8) * value / 1e9 / peak,
    }

    # ---- the same SpMV kernel timed ALONE (burst clocks: MEASURED_PEAKS' copy figure is a burst figure too), N=1 only.
    #      Inside the long solve the GPU runs under its power cap; this is the kernel's own roofline fraction.
    if world == 1 and not args.no_c2:
        xa = [1.0 + ((torch.arange(n_loc, device=dev) + 3 * j) % 17).double() / 17.0 for j in range(2)]
        ya = [torch.empty(n_loc, dtype=torch.float64, device=dev) for _ in range(2)]
        for j in range(3):
            A.mul_vec_dev(xa[j % 2].data_ptr(), ya[j % 2].data_ptr())
        torch.cuda.synchronize()
        ea0, ea1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea0.record()
        for j in range(10):
            A.mul_vec_dev(xa[j % 2].data_ptr(), ya[j % 2].data_ptr())
        ea1.record()
        torch.cuda.synchronize()
        ms_alone = ea0.elapsed_time(ea1) / 10
        roofline["alone"] = {"ms": ms_alone, "gbs": stream / (ms_alone * 1e-3) / 1e9, "frac": stream / (ms_alone * 1e-3) / 1e9 / peak,
                             "launches": 10, "note": "10 back-to-back launches of the same kernel outside the solve (x / y alternate over 2 buffer pairs; "
                                                     "1 GB of x per launch never survives in the 126 MB L2)"}
        del xa, ya

    # ---- BASELINE.json configs[1]: standalone SpMV on 256^3 7-point (N=1 only)
    spmv_c2 = None
    if world == 1 and not args.no_c2:
        del S, M, A
        torch.cuda.empty_cache()
        n1 = 256
        A2 = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, n1, n1, n1, params=(0.0,), ctx=ctx)
        n2 = n1**3
        k = torch.arange(n2, device=dev)
        nrot = 4  # x / y pairs: a launch never finds its own x (evict_last) left in L2 by the previous one
        xs2 = [1.0 + ((k + 3 * j) % 17).double() / 17.0 for j in range(nrot)]
        ys2 = [torch.empty(n2, dtype=torch.float64, device=dev) for _ in range(nrot)]
        del k
        for j in range(5):
            A2.mul_vec_dev(xs2[j % nrot].data_ptr(), ys2[j % nrot].data_ptr())
        torch.cuda.synchronize()
        reps = 48
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(reps):
            A2.mul_vec_dev(xs2[j % nrot].data_ptr(), ys2[j % nrot].data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / reps
        e0.record()
        for j in range(reps):
            A2.mul_vec_dev(xs2[0].data_ptr(), ys2[0].data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms2_same = e0.elapsed_time(e1) / reps
        b2 = spmv_bytes(n2, A2.nnz, 4)
        spmv_c2 = {"workload": "BASELINE.json configs[1]: CSR SpMV f64, 3-D 7-point 256^3", "ms": ms2, "bytes": b2,
                   "gbs": b2 / (ms2 * 1e-3) / 1e9, "frac_of_peak": b2 / (ms2 * 1e-3) / 1e9 / peak, "launches": reps,
                   "note": f"1.74 GB per launch > 126 MB L2; x / y rotate over {nrot} buffer pairs (no x survives in L2 between launches)",
                   "same_buffer_ms": ms2_same, "same_buffer_gbs": b2 / (ms2_same * 1e-3) / 1e9}
        del A2, xs2, ys2
        torch.cuda.empty_cache()

    # ---- the other BASELINE configs at full size (N=1 only; seconds): it/s + bit-for-bit parity flags
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}

        def solve_cfg(name, G, cls, Mc, ones_value, cplx, reps):
            nloc = G.n_local
            tdt = torch.complex128 if cplx else torch.float64
            o_ = torch.full((nloc,), ones_value, dtype=tdt, device=dev)
            r_ = torch.empty(nloc, dtype=tdt, device=dev)
            x_ = torch.zeros(nloc, dtype=tdt, device=dev)
            torch.cuda.synchronize()
            G.mul_vec_dev(o_.data_ptr(), r_.data_ptr())
            Sc = cls(G, nloc).record_history(16384)
            best = None
            for _ in range(reps + 1):
                x_.zero_()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = ctx.launch_count
                a0.record()
                it_, res_ = Sc.solve_dev(r_.data_ptr(), x_.data_ptr(), 10000, 1e-8, precond=Mc)
                a1.record()
                torch.cuda.synchronize()
                ms_ = a0.elapsed_time(a1)
                best = ms_ if best is None else min(best, ms_)
                nl_ = ctx.launch_count - l0
            return {"iterations": it_, "rel_residual": res_, "solve_ms": best, "iters_per_s": it_ / (best * 1e-3), "us_per_iter": 1e3 * best / max(it_, 1),
                    "launches_per_solve": nl_, "parity": gold_match(gold, name, it_, res_, Sc.history)}

        try:
            # C1: the reference's Dirichlet matrix (src/main.rs:53-88) generated on the device; rhs = i + j on the border
            g1 = 512
            A1 = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g1, g1, 1, ctx=ctx)
            ii, jj = np.meshgrid(np.arange(g1), np.arange(g1), indexing="ij")
            border = (ii == 0) | (ii == g1 - 1) | (jj == 0) | (jj == g1 - 1)
            rhs1 = torch.from_numpy(np.where(border, (ii + jj).astype(np.float64), 0.0).ravel()).to(dev)
            x1 = torch.zeros(g1 * g1, dtype=torch.float64, device=dev)
            M1 = sp.DiagPrecond.from_matrix(A1)
            S1 = sp.BiCGStab(A1, g1 * g1).record_history(16384)
            best, nl1 = None, 0
            for _ in range(4):
                x1.zero_()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0 = ctx.launch_count
                a0.record()
                it1, res1 = S1.solve_dev(rhs1.data_ptr(), x1.data_ptr(), 10000, 1e-8, precond=M1)
                a1.record()
                torch.cuda.synchronize()
                nl1 = ctx.launch_count - l0
                ms1 = a0.elapsed_time(a1)
                best = ms1 if best is None else min(best, ms1)
            configs["c1"] = {"workload": "BASELINE.json configs[0]: Jacobi-BiCGStab f64, reference 2-D 5-point Dirichlet matrix 512^2, rtol 1e-8",
                             "iterations": it1, "rel_residual": res1, "solve_ms": best, "iters_per_s": it1 / (best * 1e-3),
                             "us_per_iter": 1e3 * best / max(it1, 1), "launches_per_solve": nl1,
                             "path": "single cooperative kernel (L2-resident system)" if nl1 <= 2 else "multi-kernel loop",
                             "parity": gold_match(gold, "c1_512", it1, res1, S1.history)}
            del S1, M1, A1, rhs1, x1
            A3 = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, 128, 128, 128, params=(0.05,), ctx=ctx)
            t0 = time.perf_counter()
            M3 = sp.GaussSeidelPrecond(A3, symmetric=True)
            gs_setup = time.perf_counter() - t0
            configs["c3"] = {"workload": "BASELINE.json configs[2]: MINRES f64, shifted 7-point 128^3, symmetric Gauss-Seidel preconditioner, rtol 1e-8",
                             "gs_analysis_seconds": gs_setup} | solve_cfg("c3_128", A3, sp.MinRes, M3, 1.0, False, 1)
            configs["c3_plain"] = {"workload": "the same system without preconditioner"} | solve_cfg("c3_128_plain", A3, sp.MinRes, None, 1.0, False, 2)
            del M3, A3
            torch.cuda.empty_cache()
            A4 = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, 200, 200, 200, params=(0.5, 0.5), dtype=np.complex128, ctx=ctx)
            configs["c4"] = {"workload": "BASELINE.json configs[3]: CSMINRES complex128, complex-symmetric Helmholtz 7-point 200^3, rtol 1e-8"} | solve_cfg(
                "c4_200", A4, sp.CSMinRes, None, 1 + 1j, True, 2)
            del A4
            torch.cuda.empty_cache()
        except Exception as e:  # the headline line must survive a failure in the extras
            configs["error"] = f"{type(e).__name__}: {e}"


    # ---- CPU baseline on rank 0 at N=1
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        cpu_s = cpu_bicgstab_sample(args.cpu_grid, args.cpu_iters, g)
        cpu = {k: cpu_s[k] for k in ("value", "unit", "cores", "kind", "sample")}
        small = cpu_bicgstab_sample(192, args.cpu_iters, g)
        cpu["extrapolation_check"] = {"grid_a": 192, "value_a": small["value"], "grid_b": args.cpu_grid, "value_b": cpu_s["value"],
                                      "ratio_a_over_b": small["value"] / cpu_s["value"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world) | {"iters_per_step": iters},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n_loc * 8 * world, "d2h_bytes_per_step": n_loc * 8 * world,
                    "ms_per_step": ms_e2e / args.steps, "api": "BiCGStab.precond_solve on pinned host slices (spb_solver_solve)",
                    "max_abs_diff_vs_device_resident_step": e2e_check},
            "gpu_launches": launches, "clocks": clocks, "full_solve": full, "spmv_c2": spmv_c2, "configs": configs,
            "parity_check": parity, "cpu_binding": cpu_bind, "setup_seconds": setup_s,
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
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 extras (N=1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the exact-dot golden check before timing")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
