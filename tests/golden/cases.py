"""Golden cases: small deterministic instances of the five BASELINE.json configs plus the operator
level (SpMV, mul_vec_dot, Jacobi, Gauss-Seidel apply, the stationary GaussSeidel solver).

`build_inputs(orc, name)` regenerates a case's inputs from the oracle's generators (no RNG, or a
fixed numpy seed); `CASES` lists them.  The committed golden_v1.npz was produced by
make_golden.py with the CPU oracle (oracle/sprs_oracle.cpp), which restates the reference's
sequential code paths (src/mat.rs:96-105, src/vecalg.rs:556-605, src/bicg_stab.rs, src/minres.rs,
src/cs_minres.rs, src/gauss_seidel.rs) and is itself pinned to the reference's known-answer tests
(tests/test_oracle_kats.py).  The reference is a Rust crate that cannot be built in this image
(no cargo/rustc, nightly + MKL + unvendored git dependencies), so these vectors are the oracle's,
not the reference binary's."""
from __future__ import annotations

import numpy as np

HIST_K = 60  # residual-history entries kept per solver case

# name -> (kind, solver, preconditioner, tol, max_iter)
CASES = {
    "c1_dirichlet2d_96_jacobi_bicgstab": ("solve", "bicgstab", "diag", 1e-8, 5000),
    "c5_convdiff27_12x11x10_jacobi_bicgstab": ("solve", "bicgstab", "diag", 1e-8, 500),
    "c3_lap3d7_12_shift005_sgs_minres": ("solve", "minres", "gs_sym", 1e-8, 400),
    "c3_lap3d7_12_shift005_minres": ("solve", "minres", None, 1e-8, 400),
    "c4_helmholtz_10_csminres": ("solve", "csminres", None, 1e-8, 600),
    "gs_solver_dirichlet2d_10": ("gs_solver", None, None, 0.0, 300),
    "op_spmv_convdiff27_9x8x7": ("spmv", None, None, 0, 0),
    "op_spmv_dot_helmholtz_7x6x5": ("spmv_dot", None, None, 0, 0),
    "op_jacobi_convdiff27_9x8x7": ("jacobi", None, None, 0, 0),
    "op_gs_forward_lap3d7_9x8x7": ("gs_apply", None, "gs_fwd", 0, 0),
    "op_gs_symmetric_convdiff27_7x6x5": ("gs_apply", None, "gs_sym", 0, 0),
}


def _vec(n, dtype, phase=0.0):
    k = np.arange(n, dtype=np.float64)
    v = np.cos(0.37 * k + phase) + 0.25 * np.sin(1.3 * k)
    if np.issubdtype(np.dtype(dtype), np.complexfloating):
        v = v + 1j * np.sin(0.11 * k + 0.5 + phase)
    return v.astype(dtype)


def build_inputs(orc, name):
    """Returns (A, rhs_or_x)."""
    if name.startswith("c1_") :
        return orc.gen_dirichlet2d(96)
    if name.startswith("c5_"):
        A = orc.gen_convdiff27(12, 11, 10)
        return A, orc.spmv(A, np.ones(A.n))
    if name.startswith("c3_"):
        A = orc.gen_lap3d7(12, 12, 12, shift=0.05)
        return A, orc.spmv(A, np.ones(A.n))
    if name.startswith("c4_"):
        A = orc.gen_lap3d7(10, 10, 10, shift=0.5 + 0.5j, dtype=np.complex128)
        return A, orc.spmv(A, _vec(A.n, np.complex128))  # a structureless solution: (1+i)*ones makes the history rounding noise
    if name.startswith("gs_solver_"):
        return orc.gen_dirichlet2d(10)
    if name in ("op_spmv_convdiff27_9x8x7", "op_jacobi_convdiff27_9x8x7"):
        A = orc.gen_convdiff27(9, 8, 7)
        return A, _vec(A.n, np.float64)
    if name == "op_spmv_dot_helmholtz_7x6x5":
        A = orc.gen_lap3d7(7, 6, 5, shift=0.5 + 0.5j, dtype=np.complex128)
        return A, _vec(A.n, np.complex128)
    if name == "op_gs_forward_lap3d7_9x8x7":
        A = orc.gen_lap3d7(9, 8, 7, shift=0.05)
        return A, _vec(A.n, np.float64, 0.3)
    if name == "op_gs_symmetric_convdiff27_7x6x5":
        A = orc.gen_convdiff27(7, 6, 5)
        return A, _vec(A.n, np.float64, 0.7)
    raise KeyError(name)


def pc_of(A, pc):
    if pc is None:
        return None
    return ("diag", A.diagonal()) if pc == "diag" else (pc,)


def oracle_outputs(orc, name):
    """Runs the CPU oracle on a case; returns a dict of arrays (what golden_v1.npz stores)."""
    kind, solver, pc, tol, max_iter = CASES[name]
    A, v = build_inputs(orc, name)
    if kind == "solve":
        kw = dict(max_iter=max_iter, tol=tol, hist_cap=max_iter + 1)
        o = orc.csminres(A, v, **kw) if solver == "csminres" else getattr(orc, solver)(A, v, pc=pc_of(A, pc), **kw)
        assert o.status == orc.OK, (name, o.status)
        # How far the REFERENCE ALGORITHM ITSELF moves when only the order of its long sums changes
        # (serial fold vs OpenMP partial sums over 2/3/5/8 threads = the MKL / rayon build flavours):
        # the bound any implementation that re-orders reductions can be held to.
        k = min(HIST_K, len(o.hist))
        floor = np.zeros(k)
        its = [o.iters]
        try:
            for nt in (2, 3, 5, 8):
                orc.set_threads(nt)
                orc.set_mode(2)
                w = orc.csminres(A, v, **kw) if solver == "csminres" else getattr(orc, solver)(A, v, pc=pc_of(A, pc), **kw)
                its.append(w.iters)
                m = min(k, len(w.hist))
                floor[:m] = np.maximum(floor[:m], np.abs(w.hist[:m] - o.hist[:m]) / np.abs(o.hist[:m]))
                floor[m:] = np.inf
        finally:
            orc.set_mode(0)
            orc.set_threads(orc.max_threads())
        return {"iters": np.int64(o.iters), "resid": np.float64(o.resid), "hist": o.hist[:HIST_K].copy(), "x": o.x,
                "reorder_floor": np.maximum.accumulate(floor), "iters_range": np.array([min(its), max(its)], np.int64)}
    if kind == "gs_solver":
        o = orc.gauss_seidel(A, v, max_iter=max_iter, eps=tol)
        return {"iters": np.int64(o.iters), "resid": np.float64(o.resid), "hist": o.hist[:HIST_K].copy(), "x": o.x,
                "status": np.int64(o.status)}
    if kind == "spmv":
        return {"y": orc.spmv(A, v)}
    if kind == "spmv_dot":
        y, d = orc.spmv_dot(A, v)
        return {"y": y, "dot": np.complex128(d)}
    if kind == "jacobi":
        return {"y": orc.diag_apply(A.diagonal(), v)}
    if kind == "gs_apply":
        return {"y": orc.gs_apply(A, v, pc == "gs_sym")}
    raise KeyError(kind)
