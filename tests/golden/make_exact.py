"""Generates tests/golden/exact_v1.npz: residual histories of the BASELINE configs at FULL size from
the oracle's exact-dot flavour (oracle/sprs_oracle.cpp, mode 3: the reference's algorithms with every
dot / norm computed as the exact sum rounded once -- the one summation order that is no order).

    python tests/golden/make_exact.py [case ...]        (CPU only; minutes on 8 cores)

A second implementation whose element-wise work follows the reference operation for operation and whose
reductions are exactly rounded must reproduce these histories BIT FOR BIT over every iteration
(tests/test_gpu_exact.py).  The results do not depend on the thread count (exact sums; the
level-parallel Gauss-Seidel sweeps are bit-identical to the sequential ones), so the file can be
regenerated anywhere.  The sequential flavour (mode 0 = the reference as written, non-MKL) is run
next to it where that is affordable: its iteration count and its distance from the exact-dot history
are the reference's own rounding noise (profiles/r02_oracle_noise.md is printed from these fields).
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

OUT = os.path.join(HERE, "exact_v1.npz")


def sha(x: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


def c1(g):
    A, rhs = orc.gen_dirichlet2d(g)
    return A, rhs, "bicgstab", dict(max_iter=10000, tol=1e-8, pc=("diag", A.diagonal()))


def c3(g, pc=True):
    A = orc.gen_lap3d7(g, shift=0.05)
    rhs = orc.spmv(A, np.ones(A.n), parallel=True)
    return A, rhs, "minres", dict(max_iter=10000, tol=1e-8, pc=("gs_sym",) if pc else None)


def c4(g):
    A = orc.gen_lap3d7(g, shift=0.5 + 0.5j, dtype=np.complex128)
    rhs = orc.spmv(A, np.full(A.n, 1 + 1j), parallel=True)
    return A, rhs, "csminres", dict(max_iter=10000, tol=1e-8)


def c5(g):
    A = orc.gen_convdiff27(g)
    rhs = orc.spmv(A, np.ones(A.n), parallel=True)
    return A, rhs, "bicgstab", dict(max_iter=10000, tol=1e-8, pc=("diag", A.diagonal()))


# name -> (builder, also run the sequential flavour?)
CASES = {
    "c1_512": (lambda: c1(512), True),
    "c3_128": (lambda: c3(128), True),
    "c3_128_plain": (lambda: c3(128, pc=False), True),
    "c4_200": (lambda: c4(200), True),
    "c5_48": (lambda: c5(48), True),
    "c5_96": (lambda: c5(96), True),
    "c5_192": (lambda: c5(192), True),
}


def run(name):
    build, with_seq = CASES[name]
    A, rhs, solver, kw = build()
    out = {}
    for mode, tag in ((3, "exact"), (1, "seq")):  # mode 1 = sequential folds, row-parallel SpMV / level-parallel GS (same bits as mode 0)
        if mode == 1 and not with_seq:
            continue
        orc.set_mode(mode)
        t = time.time()
        o = getattr(orc, solver)(A, rhs, hist_cap=kw["max_iter"] + 1, **kw)
        dt = time.time() - t
        print(f"{name:14s} {tag:5s} status={o.status} iters={o.iters} resid={o.resid:.6e} hist={len(o.hist)} {dt:.1f}s", flush=True)
        out[f"{name}.{tag}.hist"] = o.hist
        out[f"{name}.{tag}.iters"] = np.int64(o.iters)
        out[f"{name}.{tag}.status"] = np.int64(o.status)
        out[f"{name}.{tag}.resid"] = np.float64(o.resid)
        if mode == 3:
            out[f"{name}.exact.x_sha256"] = np.array(sha(o.x))
            out[f"{name}.exact.x_head"] = o.x[:16].copy()
            out[f"{name}.n"] = np.int64(A.n)
            out[f"{name}.nnz"] = np.int64(A.nnz)
    orc.set_mode(0)
    return out


def main():
    names = sys.argv[1:] or list(CASES)
    data = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for nm in names:
        data.update(run(nm))
        np.savez_compressed(OUT, **data)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
