#!/usr/bin/env python
"""Times the level-scheduled Gauss-Seidel sweep (SGS apply = forward + backward sweep) on the C3
matrix.  Launch knobs come from the environment (SPB_GS_BARRIER / SPB_GS_BLOCK / SPB_GS_BACKOFF /
SPB_GS_MAXCTAS / SPB_GS_LEGACY), one process per variant."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sprsolve_b200 as sp

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "lap7"
torch.cuda.set_device(0)
ctx = sp.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
if kind == "lap7":
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, g, g, g, params=(0.05,), ctx=ctx)
else:
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25), ctx=ctx)
M = sp.GaussSeidelPrecond(A, symmetric=True)
n = A.n_local
torch.manual_seed(1)
r = torch.rand(n, dtype=torch.float64, device="cuda")
z = torch.empty_like(r)
for _ in range(3):
    M.mul_vec_dev(r.data_ptr(), z.data_ptr())
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    M.mul_vec_dev(r.data_ptr(), z.data_ptr())
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
lv = M.levels()
info = M.schedule_info(stats_blocks=0)
if os.environ.get("SPB_GS_STATS"):
    info = M.schedule_info(stats_blocks=M.schedule_info()["blocks"])
    st = info.pop("stats")
    print("stats (last sweep = backward): clocks med/max", int(np.median(st[:, 0])), int(st[:, 0].max()), "ring-wait med/max", int(np.median(st[:, 1])), int(st[:, 1].max()),
          "poll retries total", int(st[:, 2].sum()), "t0 barrier clocks med/max", int(np.median(st[:, 3])), int(st[:, 3].max()))
    for b in (0, 1, 2, 3, 8, 16, 32, 64, 96, 127, 147):
        if b < len(st):
            print("  ticket", b, "clocks", int(st[b, 0]), "ring", int(st[b, 1]), "spins", int(st[b, 2]), "bar", int(st[b, 3]))
print(info)
knobs = {k: v for k, v in os.environ.items() if k.startswith("SPB_GS")}
print(f"{kind} {g}^3 levels={lv} sgs_apply_ms={ms:.4f} us_per_level={1e3 * ms / (lv[0] + lv[1]):.3f} checksum={float(z.sum()):.17g} {knobs}")
