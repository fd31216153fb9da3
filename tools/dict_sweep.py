"""Launch-plan sweep of the dictionary SpMV on the 27-point matrix (development tool)."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sprsolve_b200 as sp
g = int(sys.argv[1]) if len(sys.argv) > 1 else 384
ctx = sp.default_context()
n = g ** 3
x = torch.ones(n, dtype=torch.float64, device="cuda") * 1.5
y = torch.empty(n, dtype=torch.float64, device="cuda")
for ct, st, bps in itertools.product((64, 96, 128, 160, 192, 224, 256), (1, 2, 3), (0,)):
    os.environ["SPB_SPMV_CT"] = str(ct); os.environ["SPB_SPMV_STAGES"] = str(st)
    if bps: os.environ["SPB_SPMV_BPS"] = str(bps)
    else: os.environ.pop("SPB_SPMV_BPS", None)
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25))
    for _ in range(3): A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    ctx.synchronize(); ctx.profile_reset(); ctx.profile(True)
    for _ in range(10): A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    nl, ms = ctx.profile_read(0); ctx.profile(False); ctx.profile_reset()
    nnz = A.nnz
    b = nnz * 12 + (n + 1) * (8 if nnz >= 2**31 - 8 else 4) + 16 * n
    print(f"ct={ct} stages={st} bps={bps}: {ms/nl:.4f} ms  {b/(ms/nl*1e-3)/1e9:.0f} GB/s algorithmic", flush=True)
    A.destroy()
