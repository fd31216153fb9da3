"""Per-phase clocks of the single-kernel BiCGStab (SPB_FUSED_STATS=1), min / median / max over the CTAs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

dev = torch.device("cuda:0")
ctx = sp.default_context()
os.environ["SPB_FUSED"] = "1"
os.environ["SPB_FUSED_STATS"] = "1"
for g in [int(v) for v in os.environ.get("TUNE_GRIDS", "100,256,512").split(",")]:
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g, g, 1)
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    border = (ii == 0) | (ii == g - 1) | (jj == 0) | (jj == g - 1)
    rhs = torch.from_numpy(np.where(border, (ii + jj).astype(np.float64), 0.0).ravel()).to(dev)
    x = torch.zeros(g * g, dtype=torch.float64, device=dev)
    M = sp.DiagPrecond.from_matrix(A)
    S = sp.BiCGStab(A, g * g)
    for block in (256, 512):
        for mode, (smem, win, cl) in {"smem": ("1", "", "0"), "cluster": ("1", "", "1")}.items():
            os.environ["SPB_FUSED_CLUSTER"] = cl
            os.environ["SPB_FUSED_BLOCK"] = str(block)
            os.environ["SPB_FUSED_SMEM"] = smem
            x.zero_()
            print(f"grid {g} block {block} {mode}:", file=sys.stderr, flush=True)
            S.solve_dev(rhs.data_ptr(), x.data_ptr(), 10000, 1e-8, precond=M)
            torch.cuda.synchronize()
