"""Default plan vs autotuned plan on the BASELINE stencils (development tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sprsolve_b200 as sp
PEAK = 6550.4
ctx = sp.default_context()
def run(name, kind, n1, params, dtype, tune):
    A = sp.GpuCsrMat.from_stencil(kind, n1, n1, n1, params=params, dtype=dtype)
    if tune:
        A.mv_hint(2000)
    n = n1 ** 3
    tdt = torch.float64 if dtype == "float64" else torch.complex128
    x = torch.ones(n, dtype=tdt, device="cuda") * 1.5
    y = torch.empty(n, dtype=tdt, device="cuda")
    for _ in range(3):
        A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    ctx.synchronize(); ctx.profile_reset(); ctx.profile(True)
    for _ in range(20):
        A.mul_vec_dev(x.data_ptr(), y.data_ptr())
    nl, ms = ctx.profile_read(0); ctx.profile(False); ctx.profile_reset()
    vb = 8 if dtype == "float64" else 16
    nnz = A.nnz
    b = nnz * (vb + 4) + (n + 1) * (8 if nnz >= 2**31 - 8 else 4) + 2 * n * vb
    print(f"{name:20s} {'autotuned' if tune else 'static   '}: {b / (ms / nl * 1e-3) / 1e9:8.1f} GB/s {100 * b / (ms / nl * 1e-3) / 1e9 / PEAK:6.1f}%  {ms / nl:.4f} ms", flush=True)
    A.destroy()
cases = [("lap7 256^3 f64", sp.STENCIL_LAP3D7, 256, (0.0,), "float64"), ("cd27 256^3 f64", sp.STENCIL_CONVDIFF27, 256, (1.0, 0.5, 0.25), "float64"),
         ("helm7 200^3 c128", sp.STENCIL_LAP3D7, 200, (0.5, 0.5), "complex128"), ("lap7 128^3 f64", sp.STENCIL_LAP3D7, 128, (0.05,), "float64"),
         ("cd27 384^3 f64", sp.STENCIL_CONVDIFF27, 384, (1.0, 0.5, 0.25), "float64"), ("cd27 512^3 f64", sp.STENCIL_CONVDIFF27, 512, (1.0, 0.5, 0.25), "float64")]
for c in cases:
    for tune in (False, True):
        if c[1] == sp.STENCIL_DIRICHLET2D:
            A = None
        run(*c, tune)
