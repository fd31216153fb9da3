"""Launch-plan sweep of the dictionary + x-window SpMV (development tool): ms per product and GB/s of the
bytes the kernel streams, window on / off, on the 27-point and the 7-point matrix."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "27"
g = int(sys.argv[2]) if len(sys.argv) > 2 else 384
ctx = sp.default_context()
n = g**3
k = torch.arange(n, device="cuda")
xs = [1.0 + ((k + 3 * j) % 17).double() / 17.0 for j in range(3)]
ys = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
del k
cts = [int(c) for c in os.environ.get("SWEEP_CT", "0,64,96,128,192,256").split(",")]
for xw, ct, st in itertools.product(("1", "0"), cts, (1, 2)):
    os.environ["SPB_SPMV_XWIN"] = xw
    if ct:
        os.environ["SPB_SPMV_CT"] = str(ct)
        os.environ["SPB_SPMV_STAGES"] = str(st)
    else:
        os.environ.pop("SPB_SPMV_CT", None)
        os.environ.pop("SPB_SPMV_STAGES", None)
        if st == 2:
            continue
    if kind == "27":
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25))
    else:
        A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.0,))
    pi = A.plan_info()
    for j in range(3):
        A.mul_vec_dev(xs[j].data_ptr(), ys[j].data_ptr())
    ctx.synchronize()
    ctx.profile_reset()
    ctx.profile(True)
    for j in range(12):
        A.mul_vec_dev(xs[j % 3].data_ptr(), ys[j % 3].data_ptr())
    nl, ms = ctx.profile_read(0)
    ctx.profile(False)
    ctx.profile_reset()
    t = ms / nl
    print(f"{kind}-pt {g}^3 xwin={xw} ct={ct or 'auto'} stages={st if ct else 'auto'} -> dict={pi['dictionary']} win={pi['x_window']} "
          f"plan ct={pi['consumer_threads']} st={pi['stages']} tile={pi['tile_nnz']} bps={pi['ctas_per_sm']}: {t:.4f} ms  "
          f"{pi['stream_bytes'] / (t * 1e-3) / 1e9:.0f} GB/s streamed", flush=True)
    A.destroy()
