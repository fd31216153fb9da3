#!/usr/bin/env python
"""Randomised stress of the block-wavefront Gauss-Seidel sweep against the sequential oracle:
random sparse patterns (banded + long-range entries), sizes, block / stage geometries, both scalar
types, forward / symmetric applies and the stationary solver's sweep.  Development tool:
    python tools/gs_stress.py [cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import sprsolve_b200 as sp
from oracle import oracle as orc

orc.build()
orc.set_mode(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = sp.default_context()


def random_matrix(n, dtype):
    band = int(rng.integers(1, 40))
    per_row = int(rng.integers(1, 12))
    far = rng.random() < 0.5
    ip, idx, val = [0], [], []
    for i in range(n):
        c = i + rng.integers(-band, band + 1, size=per_row)
        if far:
            c = np.concatenate([c, rng.integers(0, n, size=int(rng.integers(0, 3)))])
        c = np.unique(np.concatenate([np.clip(c, 0, n - 1), [i]]))
        v = rng.uniform(-1, 1, size=c.size).astype(dtype)
        if np.issubdtype(dtype, np.complexfloating):
            v = v + 1j * rng.uniform(-1, 1, size=c.size)
        v[c == i] = c.size + 1.0
        idx.append(c)
        val.append(v)
        ip.append(ip[-1] + c.size)
    return orc.Csr(n, np.array(ip, np.int64), np.concatenate(idx).astype(np.int32), np.concatenate(val))


bad = 0
for case in range(cases):
    dtype = np.complex128 if rng.random() < 0.3 else np.float64
    n = int(rng.integers(2, 4000))
    knobs = {}
    if rng.random() < 0.8:
        knobs["SPB_GS_BLOCK_ROWS"] = str(int(rng.integers(1, max(2, n))))
    if rng.random() < 0.6:
        knobs["SPB_GS_STAGE_ROWS"] = str(int(rng.integers(8, 300)))
        knobs["SPB_GS_STAGE_BYTES"] = str(int(rng.integers(2048, 40000)))
        knobs["SPB_GS_STAGE_OTHER"] = str(int(rng.integers(64, 3000)))
    if rng.random() < 0.5:
        knobs["SPB_GS_STAGES"] = str(int(rng.integers(2, 6)))
    for k in list(os.environ):
        if k.startswith("SPB_GS_"):
            del os.environ[k]
    os.environ.update(knobs)
    A = random_matrix(n, dtype)
    G = sp.GpuCsrMat.new(A.indptr, A.indices, A.data, ctx=ctx)
    ok = True
    info = None
    for symmetric in (False, True):
        P = sp.GaussSeidelPrecond(G, symmetric=symmetric)
        info = P.schedule_info()
        for _ in range(2):
            v = rng.uniform(-1, 1, n).astype(dtype)
            if dtype is np.complex128:
                v = v + 1j * rng.uniform(-1, 1, n)
            out = np.zeros(n, dtype)
            P.mul_vec(v, out)
            ok &= np.array_equal(out, orc.gs_apply(A, v, symmetric))
        ok &= P.schedule_info()["poll_timeout"] == 0
    if dtype is np.float64:
        rhs = orc.spmv(A, np.ones(n))
        x = np.zeros(n)
        o = orc.gauss_seidel(A, rhs, max_iter=6, eps=0.0)
        try:
            sp.GaussSeidel(G).solve(rhs, x, 6, 0.0)
        except sp.SolverError:
            pass
        ok &= np.array_equal(x, o.x)
    if not ok:
        bad += 1
        print("MISMATCH", case, n, dtype.__name__, knobs, info, flush=True)
print(f"gs_stress: {cases} cases, {bad} mismatches (wave schedules used: see fwd_ok in the last info {info})")
sys.exit(1 if bad else 0)
