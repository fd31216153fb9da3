"""Single-kernel BiCGStab (csrc/bicgstab.cu: bicg_fused_kernel): us per iteration on the reference Dirichlet
matrix for the three modes (global vectors / shared-memory vectors / one thread-block cluster) and two CTA sizes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

dev = torch.device("cuda:0")
ctx = sp.default_context()
grids = [int(g) for g in os.environ.get("TUNE_GRIDS", "100,256,512").split(",")]
for g in grids:
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g, g, 1)
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    border = (ii == 0) | (ii == g - 1) | (jj == 0) | (jj == g - 1)
    rhs = torch.from_numpy(np.where(border, (ii + jj).astype(np.float64), 0.0).ravel()).to(dev)
    x = torch.zeros(g * g, dtype=torch.float64, device=dev)
    M = sp.DiagPrecond.from_matrix(A)
    S = sp.BiCGStab(A, g * g)
    ref = None
    for block in (256, 512):
        for mode, (smem, win, cl) in {"global": ("0", "", "0"), "smem": ("1", "", "0"), "cluster": ("1", "", "1")}.items():
            os.environ["SPB_FUSED"] = "1"
            os.environ["SPB_FUSED_CLUSTER"] = cl
            os.environ["SPB_FUSED_BLOCK"] = str(block)
            os.environ["SPB_FUSED_SMEM"] = smem
            ts = []
            for _ in range(4):
                x.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                it, res = S.solve_dev(rhs.data_ptr(), x.data_ptr(), 10000, 1e-8, precond=M)
                ts.append(time.perf_counter() - t0)
            t = min(ts[1:])
            xs = x.cpu().numpy().copy()
            if ref is None:
                ref = (it, res, xs)
            same = it == ref[0] and res == ref[1] and np.array_equal(xs, ref[2])
            print(f"grid {g:4d}  block {block:4d}  {mode:7s} its {it:5d}  solve {1e3 * t:8.3f} ms  {1e6 * t / it:7.2f} us/iter  same bits {same}", flush=True)
    if os.environ.get("TUNE_STATS"):
        os.environ["SPB_FUSED_STATS"] = "1"
        os.environ["SPB_FUSED_BLOCK"] = "512"
        os.environ["SPB_FUSED_SMEM"] = "1"
        x.zero_()
        S.solve_dev(rhs.data_ptr(), x.data_ptr(), 10000, 1e-8, precond=M)
        os.environ.pop("SPB_FUSED_STATS")
