"""Relaxed Gauss-Seidel: SOR sweep / SSOR(omega) operator and the SOR stationary solver (SURVEY.md
section 8f rank 3: "relaxed / SSOR(omega) variants"; north star: "level-scheduled Gauss-Seidel/SSOR sweep
that preserves the reference's sequential update order").

The reference's GaussSeidel (src/gauss_seidel.rs) has no relaxation factor, so the update is defined by the
oracle (oracle/sprs_oracle.h): x_i <- (1 - w) x_i + w g_i, g_i the :123 value, two real-by-scalar products
and one add, in natural row order.  Bars: the operator and the solver are BIT-EXACT against the sequential
oracle; solver histories bit for bit against its exact-dot flavour; w = 1 is the un-relaxed path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import sprsolve_b200 as s

    s.default_context()
    return s


def to_gpu(sp, A):
    return sp.GpuCsrMat.new(A.indptr, A.indices, A.data, shape=(A.n, A.ncols))


def _mats(orc):
    import scipy.sparse as sps

    M = (sps.random(900, 900, 0.006, random_state=11, format="csr") + sps.diags(np.full(900, 3.5))).tocsr()
    M.sort_indices()
    return [orc.gen_lap3d7(13, 11, 9, shift=0.05), orc.gen_convdiff27(9, 8, 7), orc.gen_dirichlet2d(14)[0],
            orc.Csr(900, M.indptr, M.indices, M.data), orc.gen_lap3d7(8, 7, 6, shift=0.5 + 0.5j, dtype=np.complex128),
            orc.Csr(900, M.indptr, M.indices, M.data.astype(np.float32))]


@pytest.mark.parametrize("omega", [0.6, 1.0, 1.4, 1.85])
def test_sor_ssor_apply_bit_exact(sp, orc, omega):
    rng = np.random.default_rng(5)
    for A in _mats(orc):
        v = rng.uniform(-1, 1, A.n).astype(A.dtype)
        if np.iscomplexobj(v):
            v = (v + 1j * rng.uniform(-1, 1, A.n)).astype(A.dtype)
        G = to_gpu(sp, A)
        for sym in (False, True):
            M = sp.GaussSeidelPrecond(G, symmetric=sym, omega=omega)
            z = np.zeros(A.n, A.dtype)
            M.mul_vec(v, z)
            assert np.array_equal(z, orc.gs_apply(A, v, sym, omega)), (omega, sym, A.dtype)
            if omega == 1.0:
                assert np.array_equal(z, orc.gs_apply(A, v, sym))


def test_ssor_is_the_textbook_operator(sp, orc):
    """z = w (2 - w) (D + w U)^-1 D (D + w L)^-1 r  (symmetric positive definite for SPD-diagonal symmetric A)."""
    A = orc.gen_lap3d7(7, 6, 5, shift=0.05)
    Md = A.to_scipy().toarray()
    D, L, U = np.diag(np.diag(Md)), np.tril(Md, -1), np.triu(Md, 1)
    v = np.cos(0.3 * np.arange(A.n))
    G = to_gpu(sp, A)
    for w in (0.8, 1.5):
        z = np.zeros(A.n)
        sp.GaussSeidelPrecond(G, symmetric=True, omega=w).mul_vec(v, z)
        ref = w * (2 - w) * np.linalg.solve(D + w * U, D @ np.linalg.solve(D + w * L, v))
        assert np.allclose(z, ref, rtol=1e-12, atol=1e-13)
    with pytest.raises(sp.BackendError):
        sp.GaussSeidelPrecond(G, symmetric=True, omega=2.0)


def test_solvers_with_ssor_bit_for_bit(sp, orc):
    """MINRES with SSOR(w) (valid: SPD), BiCGStab with a SOR sweep, the SOR stationary solver -- against the
    exact-dot oracle, every iteration."""
    try:
        orc.set_mode(3)
        A = orc.gen_lap3d7(20, 18, 16, shift=0.05)
        rhs = orc.spmv(A, np.ones(A.n))
        G = to_gpu(sp, A)
        for w in (1.2, 0.9):
            o = orc.minres(A, rhs, max_iter=600, tol=1e-8, pc=("ssor", w), hist_cap=601)
            S = sp.MinRes(G, A.n).record_history(601)
            x = np.zeros(A.n)
            it, res = S.precond_solve(sp.GaussSeidelPrecond(G, symmetric=True, omega=w), rhs, x, 600, 1e-8)
            assert o.status == 0 and (it, res) == (o.iters, o.resid)
            assert np.array_equal(S.history, o.hist) and np.array_equal(x, o.x)
        B = orc.gen_convdiff27(12, 11, 10)
        rb = orc.spmv(B, np.ones(B.n))
        GB = to_gpu(sp, B)
        o = orc.bicgstab(B, rb, max_iter=300, tol=1e-8, pc=("sor_fwd", 1.3), hist_cap=301)
        S = sp.BiCGStab(GB, B.n).record_history(301)
        x = np.zeros(B.n)
        it, res = S.precond_solve(sp.GaussSeidelPrecond(GB, symmetric=False, omega=1.3), rb, x, 300, 1e-8)
        assert o.status == 0 and (it, res) == (o.iters, o.resid) and np.array_equal(S.history, o.hist) and np.array_equal(x, o.x)
        # SOR solver on the reference's fixture matrix (tests/test_solvers.rs:3-31): far fewer sweeps than omega = 1
        C, rc = orc.gen_dirichlet2d(12)
        GC = to_gpu(sp, C)
        its = {}
        for w in (1.0, 1.5):
            o = orc.gauss_seidel(C, rc, max_iter=2000, eps=1e-9, omega=w, hist_cap=2000)
            S = sp.GaussSeidel(GC, omega=w).record_history(2000)
            x = np.zeros(C.n)
            it, res = S.solve(rc, x, 2000, 1e-9)
            assert o.status == 0 and (it, res) == (o.iters, o.resid) and np.array_equal(S.history, o.hist) and np.array_equal(x, o.x)
            its[w] = it
        assert its[1.5] < its[1.0] // 2
    finally:
        orc.set_mode(0)
