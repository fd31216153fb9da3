"""Launch-geometry sweep of the single-kernel BiCGStab (csrc/bicgstab.cu: bicg_fused_kernel) on the
reference Dirichlet matrix: us per iteration for CTA size x CTAs per SM at several grid sizes."""
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
for g in (100, 256, 512):
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g, g, 1)
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    border = (ii == 0) | (ii == g - 1) | (jj == 0) | (jj == g - 1)
    rhs = torch.from_numpy(np.where(border, (ii + jj).astype(np.float64), 0.0).ravel()).to(dev)
    x = torch.zeros(g * g, dtype=torch.float64, device=dev)
    M = sp.DiagPrecond.from_matrix(A)
    S = sp.BiCGStab(A, g * g)
    for block in [int(b) for b in os.environ.get("TUNE_BLOCKS", "256,512,1024").split(",")]:
        for smem in ("1", "0"):
            os.environ["SPB_FUSED"] = "1"
            os.environ["SPB_FUSED_BLOCK"] = str(block)
            os.environ["SPB_FUSED_SMEM"] = smem
            ts = []
            for _ in range(3):
                x.zero_()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                it, res = S.solve_dev(rhs.data_ptr(), x.data_ptr(), 10000, 1e-8, precond=M)
                ts.append(time.perf_counter() - t0)
            t = min(ts[1:])
            print(f"grid {g:4d}  block {block:5d}  smem-vectors {smem}  its {it:5d}  solve {1e3 * t:8.3f} ms  {1e6 * t / it:7.2f} us/iter", flush=True)
