"""Short single-kernel workloads for `ncu --set full` (one target per invocation, a handful of launches).

    python tools/ncu_targets.py c5      27-point 512^3 f64 (dictionary stream), 3 SpMV launches
    python tools/ncu_targets.py c2      7-point 256^3 f64 (plain CSR stream)
    python tools/ncu_targets.py c4      7-point 200^3 complex128 (plain CSR stream)
    python tools/ncu_targets.py c5f32   27-point 384^3 f32 (dictionary stream, 4 bytes per non-zero)
    python tools/ncu_targets.py k3      Jacobi-BiCGStab on 27-point 384^3, 4 iterations (bicg_k1 / k2 / k3)
    python tools/ncu_targets.py gs      symmetric Gauss-Seidel apply on 7-point 128^3 (gs_wave_kernel; use --replay-mode application)
    python tools/ncu_targets.py fused   single-kernel BiCGStab on the 100^2 reference matrix (cluster mode)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

what = sys.argv[1]
ctx = sp.default_context()
dev = torch.device("cuda:0")


def spmv(A, n, tdt, reps=3):
    k = torch.arange(n, device=dev)
    x = (1.0 + (k % 17).double() / 17.0).to(tdt)
    y = torch.empty(n, dtype=tdt, device=dev)
    del k
    torch.cuda.synchronize()
    for _ in range(reps):
        A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    print(what, A.plan_info(), float(y.abs().sum().real))


if what == "c5":
    g = 512
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25)), g**3, torch.float64)
elif what == "c2":
    g = 256
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.0,)), g**3, torch.float64)
elif what == "c4":
    g = 200
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.5, 0.5), dtype=np.complex128), g**3, torch.complex128)
elif what == "c5xw":  # the x-window variant of the 27-point kernel (opt-in), 384^3
    os.environ["SPB_SPMV_XWIN"] = "1"
    g = 384
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25)), g**3, torch.float64)
elif what == "c5g384":  # the default (gather) variant at the same size
    g = 384
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25)), g**3, torch.float64)
elif what == "c5f32":
    g = 384
    spmv(sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25), dtype=np.float32), g**3, torch.float32)
elif what == "k3":
    g = 384
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25))
    n = g**3
    ones = torch.ones(n, dtype=torch.float64, device=dev)
    rhs = torch.empty(n, dtype=torch.float64, device=dev)
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    A.mul_vec_dev(ones.data_ptr(), rhs.data_ptr())
    S = sp.BiCGStab(A, n)
    try:
        S.solve_dev(rhs.data_ptr(), x.data_ptr(), 5, 1e-30, precond=sp.DiagPrecond.from_matrix(A))
    except sp.InsufficientIterNum:
        pass
    ctx.synchronize()
    print("k3 done")

if what == "gs":
    g = 128
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.05,))
    M = sp.GaussSeidelPrecond(A, symmetric=True)
    r = torch.rand(g**3, dtype=torch.float64, device=dev)
    z = torch.empty_like(r)
    M.mul_vec_dev(r.data_ptr(), z.data_ptr())
    ctx.synchronize()
    print(what, M.schedule_info(), float(z.sum()))
elif what == "fused":
    g = 100
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_DIRICHLET2D, g, g, 1)
    ii, jj = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    border = (ii == 0) | (ii == g - 1) | (jj == 0) | (jj == g - 1)
    rhs = torch.from_numpy(np.where(border, (ii + jj).astype(np.float64), 0.0).ravel()).to(dev)
    x = torch.zeros(g * g, dtype=torch.float64, device=dev)
    M = sp.DiagPrecond.from_matrix(A)
    S = sp.BiCGStab(A, g * g)
    print(what, S.solve_dev(rhs.data_ptr(), x.data_ptr(), 10000, 1e-8, precond=M))
