"""Fixed cost per launch of the dictionary SpMV: 27-point matrices 512 x 512 x nz for a range of nz, time per launch
(rotating x / y pairs) -> least-squares fit  t = F + nz * s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sprsolve_b200 as sp  # noqa: E402

ctx = sp.default_context()
dev = torch.device("cuda:0")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # the events below are recorded on torch's stream
res = []
for nz in (4, 8, 16, 32, 64, 128, 256):
    A = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, 512, 512, nz, params=(1.0, 0.5, 0.25))
    n = 512 * 512 * nz
    xs = [torch.rand(n, dtype=torch.float64, device=dev) for _ in range(4)]
    ys = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    for j in range(8):
        A.mul_vec_dev(xs[j % 4].data_ptr(), ys[j % 4].data_ptr())
    torch.cuda.synchronize()
    reps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(reps):
        A.mul_vec_dev(xs[j % 4].data_ptr(), ys[j % 4].data_ptr())
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    info = A.plan_info()
    res.append((nz, us))
    print(f"nz {nz:4d}  rows {n:10d}  {us:9.2f} us / launch   {info['stream_bytes'] / us / 1e3:7.0f} GB/s   plan ct {info['consumer_threads']} stages {info['stages']} ctas/SM {info['ctas_per_sm']}", flush=True)
    A.destroy()
    del xs, ys
a = np.array(res, dtype=float)
s, F = np.polyfit(a[:, 0], a[:, 1], 1)
print(f"fit: t = {F:.1f} us + nz * {s:.3f} us   (nz = 64 is the per-rank slab of the 512^3 system on 8 GPUs)")
