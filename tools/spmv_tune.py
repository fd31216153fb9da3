"""SpMV kernel-configuration sweep (development tool): achieved algorithmic GB/s per config on
the BASELINE stencils.  Usage: python tools/spmv_tune.py [cfg ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sprsolve_b200 as sp

PEAK = 6550.4


def bench(kind, n1, params, dtype, plan, reps=20):
    for k in list(os.environ):
        if k.startswith("SPB_SPMV_"):
            del os.environ[k]
    os.environ.update(plan)
    ctx = sp.default_context()
    A = sp.GpuCsrMat.from_stencil(kind, n1, n1, n1, params=params, dtype=dtype)
    n = n1**3
    tdt = torch.float64 if dtype == "float64" else torch.complex128
    x = torch.ones(n, dtype=tdt, device="cuda") * 1.5
    y = torch.empty(n, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    ctx.profile_reset()
    ctx.profile(True)
    for _ in range(reps):
        A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    nl, ms = ctx.profile_read(0)
    ctx.profile(False)
    ctx.profile_reset()
    vb = 8 if dtype == "float64" else 16
    nnz = A.nnz
    b = nnz * (vb + 4) + (n + 1) * (8 if nnz >= 2**31 - 8 else 4) + 2 * n * vb
    gbs = b / (ms / nl * 1e-3) / 1e9
    A.destroy()
    return gbs, ms / nl


if __name__ == "__main__":
    import itertools

    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    cases = [
        ("lap7 256^3 f64", sp.STENCIL_LAP3D7, 256, (0.0,), "float64"),
        ("cd27 256^3 f64", sp.STENCIL_CONVDIFF27, 256, (1.0, 0.5, 0.25), "float64"),
        ("helm7 200^3 c128", sp.STENCIL_LAP3D7, 200, (0.5, 0.5), "complex128"),
    ]
    plans = [{}]
    for ax, ct, st in itertools.product((0, 1), (32, 64, 96, 128, 192, 256), (1, 2)):
        plans.append({"SPB_SPMV_ASYNCX": str(ax), "SPB_SPMV_CT": str(ct), "SPB_SPMV_STAGES": str(st)})
    for name, kind, n1, params, dt in cases:
        if which != "all" and which not in name:
            continue
        for plan in plans:
            tag = " ".join(f"{k[9:]}={v}" for k, v in plan.items()) or "default"
            try:
                gbs, ms = bench(kind, n1, params, dt, plan)
                print(f"{name:18s} {tag:32s}: {gbs:8.1f} GB/s  {100 * gbs / PEAK:5.1f}%  {ms:.4f} ms", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name:18s} {tag:32s}: FAILED {e}", flush=True)
