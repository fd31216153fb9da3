"""Diagnostic (not a test): GPU-vs-oracle residual-history deviation next to the oracle's own
serial-vs-OpenMP (summation order) spread."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import sprsolve_b200 as sp
from oracle import oracle as o
import fixtures as fx

def run(name, A, rhs, solver, pc, tol=1e-8, mi=3000):
    kw = dict(max_iter=mi, tol=tol, hist_cap=mi + 1)
    f = getattr(o, solver)
    o.set_mode(0); a = f(A, rhs, **kw) if solver == "csminres" else f(A, rhs, pc=pc, **kw)
    o.set_mode(2); b = f(A, rhs, **kw) if solver == "csminres" else f(A, rhs, pc=pc, **kw)
    o.set_mode(0)
    G = sp.GpuCsrMat.new(A.indptr, A.indices, A.data)
    S = {"bicgstab": sp.BiCGStab, "minres": sp.MinRes, "csminres": sp.CSMinRes}[solver](G, A.n).record_history(mi + 1)
    x = np.zeros(A.n, A.dtype)
    P = None
    if pc is not None:
        P = sp.DiagPrecond.new(pc[1], dtype=A.dtype) if pc[0] == "diag" else sp.GaussSeidelPrecond(G, pc[0] == "gs_sym")
    try:
        it, res = S.solve(rhs, x, mi, tol) if P is None else S.precond_solve(P, rhs, x, mi, tol)
    except sp.SolverError as e:
        it, res = -1, repr(e)
    h = S.history
    m = min(50, len(a.hist), len(h), len(b.hist))
    dg = np.abs(h[:m] - a.hist[:m]) / np.abs(a.hist[:m])
    dc = np.abs(b.hist[:m] - a.hist[:m]) / np.abs(a.hist[:m])
    print(f"{name:28s} its gpu/orc/orc_omp {it}/{a.iters}/{b.iters}  max50 gpu {dg.max():.2e} cpu-reorder {dc.max():.2e}  @10 {dg[:10].max():.1e}/{dc[:10].max():.1e} @25 {dg[:25].max():.1e}/{dc[:25].max():.1e}")

for n in (48, 96, 192):
    A, rhs = o.gen_dirichlet2d(n)
    run(f"dirichlet{n} jacobi", A, rhs, "bicgstab", ("diag", A.diagonal()))
    run(f"dirichlet{n} nopc", A, rhs, "bicgstab", None)
A = o.gen_convdiff27(24); rhs = o.spmv(A, np.ones(A.n))
run("cd27 24^3 jacobi", A, rhs, "bicgstab", ("diag", A.diagonal()))
rng = np.random.default_rng(12345); rhs2 = rng.uniform(-1, 1, A.n)
run("cd27 24^3 jacobi rand rhs", A, rhs2, "bicgstab", ("diag", A.diagonal()))
A = o.gen_lap3d7(24, shift=0.05); rhs = o.spmv(A, np.ones(A.n))
run("minres lap 24^3", A, rhs, "minres", None)
run("minres lap 24^3 sgs", A, rhs, "minres", ("gs_sym",))
A = o.gen_lap3d7(20, shift=0.5 + 0.5j, dtype=np.complex128); rhs = o.spmv(A, np.full(A.n, 1 + 1j))
run("csminres helm 20^3", A, rhs, "csminres", None)
A, rhs, d, xs = fx.complex_symmetric_grid(8, 8)
run("csminres fixture", A, rhs, "csminres", None, tol=1e-12)
run("bicg csym fixture", A, rhs, "bicgstab", ("diag", d), tol=1e-12)
A, rhs, d, xs = fx.hermitian_grid(8, 8)
run("minres herm fixture", A, rhs, "minres", None, tol=1e-12)
