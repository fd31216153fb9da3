#!/usr/bin/env python
"""Short C3 run for the ncu launch list: 128^3 shifted 7-point matrix, SGS-preconditioned MINRES,
40 iterations (not converged on purpose -- the launch mix per iteration is what is profiled)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import sprsolve_b200 as sp

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
its = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.cuda.set_device(0)
ctx = sp.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.05,), ctx=ctx)
M = sp.GaussSeidelPrecond(A, symmetric=True)
n = A.n_local
ones = torch.ones(n, dtype=torch.float64, device="cuda")
rhs = torch.empty_like(ones)
A.mul_vec_dev(ones.data_ptr(), rhs.data_ptr())
x = torch.zeros_like(rhs)
S = sp.MinRes(A, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
try:
    out = S.solve_dev(rhs.data_ptr(), x.data_ptr(), its, 1e-8, precond=M)
except sp.SolverError as e:
    out = type(e).__name__
e1.record()
torch.cuda.synchronize()
print(f"C3 {g}^3 SGS-MINRES {its} iterations: {e0.elapsed_time(e1):.3f} ms ({e0.elapsed_time(e1) / its * 1e3:.1f} us/iteration) -> {out}; launches {ctx.launch_count}")
