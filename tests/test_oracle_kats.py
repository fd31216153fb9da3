"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md section 8c).  CPU only.  Reference paths are relative to the reference crate root."""
import numpy as np
import pytest

import fixtures as fx


# ---------------------------------------------------------------- SpMV KATs
def test_dense_csr_mat(orc):
    """src/mat.rs:232-255 and :258-280 (i32 indices), eps 1e-8, includes an empty row."""
    y = orc.spmv(fx.kat_csr(), np.array(fx.KAT_X))
    assert np.all(np.abs(y - np.array(fx.KAT_Y)) < 1e-8)
    assert y[1] == 0.0  # empty row is zero-filled (mat.rs:71)


def test_dense_csr_mat_parallel_same_bits(orc):
    """rayon path (mat.rs:85-107) is numerically identical to the serial one."""
    A = orc.gen_lap3d7(12, 9, 7, shift=0.3)
    x = np.linspace(-1, 2, A.n)
    assert np.array_equal(orc.spmv(A, x), orc.spmv(A, x, parallel=True))


def test_dense_csc_mat(orc):
    """src/mat.rs:208-229."""
    y = orc.spmv_csc(5, 5, fx.KAT_CSC_INDPTR, fx.KAT_CSC_INDICES, fx.KAT_CSC_DATA, fx.KAT_X)
    assert np.all(np.abs(y - np.array(fx.KAT_CSC_Y)) < 1e-8)


def test_mkl_mat_vec_complex(orc):
    """src/mkl_mat.rs:368-405: values v+vi, real x => re == im == expected."""
    y = orc.spmv(fx.kat_csr(np.complex128), np.array(fx.KAT_X, np.complex128))
    assert np.all(np.abs(y.real - np.array(fx.KAT_Y)) < 1e-8)
    assert np.all(np.abs(y.imag - np.array(fx.KAT_Y)) < 1e-8)


def test_mkl_mat_vec_2(orc):
    """src/mkl_mat.rs:408-430: integer-valued, exact to 1e-16."""
    y = orc.spmv(fx.kat2_csr(), np.array(fx.KAT2_X))
    assert np.all(np.abs(y - np.array(fx.KAT2_Y)) < 1e-16)


def test_mkl_mat_vec_dot_complex(orc):
    """src/mkl_mat.rs:433-463: dotmv == conj_dot(x, A x)."""
    A = fx.kat_csr(np.complex128)
    x = np.array(fx.KAT_X, np.complex128)
    y, d = orc.spmv_dot(A, x)
    e = orc.conj_dot(x, y)
    assert d == e
    assert abs(d - np.vdot(x, y)) < 1e-15


# ---------------------------------------------------------------- vecalg KATs (src/vecalg.rs:612-842)
def test_norm2(orc):
    assert orc.norm2(np.ones(25)) == pytest.approx(5.0, abs=1e-15)
    assert orc.norm2(np.ones(100)) == pytest.approx(10.0, abs=1e-15)
    assert orc.norm2(np.full(50, 1 + 1j)) == pytest.approx(10.0, abs=1e-14)


def test_dot_generic(orc):
    assert orc.dot(np.ones(6), np.arange(1.0, 7.0)) == 21.0


def test_conj_dot(orc):
    assert orc.conj_dot(np.full(100, 1.0), np.full(100, 2.0)) == 200.0
    # c64 doctest vecalg.rs:36-46
    a, b = np.full(100, 4 + 3j), np.full(100, 2 - 3j)
    t = np.conj(a[0]) * b[0] * 100.0
    r = orc.conj_dot(a, b)
    assert r.real == pytest.approx(t.real, abs=1e-12) and r.imag == pytest.approx(t.imag, abs=1e-12)
    # c32 flavour vecalg.rs:651-656
    import ctypes as C

    a32 = np.tile(np.array([2, 3], np.float32), 100)
    b32 = np.tile(np.array([2, -3], np.float32), 100)
    out = np.zeros(2, np.float32)
    orc.lib().orc_conj_dot_c(C.c_int64(100), a32.ctypes.data_as(C.c_void_p), b32.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    t = np.conj(np.complex64(2 + 3j)) * np.complex64(2 - 3j) * 100
    assert out[0] == pytest.approx(t.real) and out[1] == pytest.approx(t.imag)


def test_dot_no_conj(orc):
    """vecalg.rs:677-720: dot takes no conjugate: (2+3i)(2-3i) x100 = 1300 + 0i."""
    r = orc.dot(np.full(100, 2 + 3j), np.full(100, 2 - 3j))
    assert r.real == pytest.approx(1300.0) and r.imag == pytest.approx(0.0)
    r = orc.dot(np.full(100, 2 + 1j), np.full(100, 3 + 1j))
    t = (2 + 1j) * (3 + 1j) * 100
    assert r.real == pytest.approx(t.real) and r.imag == pytest.approx(t.imag)
    assert orc.dot(np.full(100, 1.0), np.full(100, 2.0)) == 200.0


def test_dot_generic_complex(orc):
    """vecalg.rs:723-747."""
    r = orc.dot(np.full(6, 1j), np.arange(6).astype(np.complex128))
    assert r.real == pytest.approx(0.0) and r.imag == pytest.approx(15.0)
    a = np.array([complex(i, i) for i in range(8)])
    b = np.array([complex(i, -i) for i in range(8)])
    r = orc.dot(a, b)
    assert r.real == pytest.approx(sum(2 * i * i for i in range(8))) and r.imag == pytest.approx(0.0)


def test_axpy_generic_complex(orc):
    """vecalg.rs:750-772."""
    a = np.full(6, 1j)
    b = np.arange(6).astype(np.complex128)
    orc.axpy(1.0, a, b)
    assert np.allclose(b.real, np.arange(6)) and np.allclose(b.imag, 1.0)
    b = np.arange(6).astype(np.complex128)
    orc.axpy(1j, a, b)
    assert np.allclose(b.real, np.arange(6) - 1.0) and np.allclose(b.imag, 0.0)


def test_axpy_axpby_f32(orc):
    """vecalg.rs:775-801 and doctest :122-132 (f32 replay)."""
    import ctypes as C

    L = orc.lib()
    a = np.ones(128, np.float32)
    b = np.zeros(128, np.float32)
    pa, pb = a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)
    for _ in range(4):
        L.orc_axpy_s(C.c_int64(128), C.c_float(2.0), pa, pb)
    assert np.all(b == 8.0)
    b[:] = 0
    for _ in range(4):
        L.orc_axpby_s(C.c_int64(128), C.c_float(2.0), pa, C.c_float(1.0), pb)
    assert np.all(b == 8.0)
    for _ in range(3):
        L.orc_axpby_s(C.c_int64(128), C.c_float(2.0), pa, C.c_float(-1.0), pb)
    assert np.all(b == -6.0)
    b[:] = 2
    L.orc_axpby_s(C.c_int64(128), C.c_float(2.0), pa, C.c_float(-1.0), pb)
    assert np.all(b == 0.0)


def test_scale_rscale_conj(orc):
    """vecalg.rs:659-674 (scale), :832-841 (rscale), :803-830 (conj)."""
    a = np.ones(100)
    orc.scale(1.5, a)
    assert np.all(a == 1.5)
    a = np.full(100, 2 + 3j)
    s = 1.2 + 4.8j
    orc.scale(s, a)
    assert np.allclose(a, (2 + 3j) * s)
    b = np.full(100, 1 + 2j)
    orc.rscale(9.0, b)
    assert np.all(b == 9 + 18j)
    c = orc.conj(np.full(100, 3 + 2j))
    assert np.all(c == 3 - 2j)


# ---------------------------------------------------------------- integration fixtures (reference tests/*.rs)
def test_gauss_seidel_fixture(orc):
    """tests/test_solvers.rs:3-31: 10x10 Dirichlet grid, solve(rhs,x,300,0.) must converge."""
    A, rhs = orc.gen_dirichlet2d(10)
    r = orc.gauss_seidel(A, rhs, max_iter=300, eps=0.0)
    assert r.status == orc.OK
    # exact solution of the discrete problem is the harmonic extension of i+j, which is i+j itself
    ii, jj = np.meshgrid(np.arange(10), np.arange(10), indexing="ij")
    assert np.allclose(r.x, (ii + jj).ravel(), atol=1e-12)


def test_bicgstab_fixture(orc):
    """tests/test_solvers.rs:34-57: 20x20, tol 1e-17, 1500 its must converge."""
    A, rhs = orc.gen_dirichlet2d(20)
    r = orc.bicgstab(A, rhs, max_iter=1500, tol=1e-17)
    assert r.status == orc.OK and r.resid <= 1e-17
    ii, jj = np.meshgrid(np.arange(20), np.arange(20), indexing="ij")
    assert np.allclose(r.x, (ii + jj).ravel(), atol=1e-11)


def test_minres_fixtures(orc):
    """tests/test_minres.rs:2-31 (symmetric 8x8) and :34-60 (diagonal)."""
    A, rhs = fx.sym_laplacian_2d(8, 8)
    S = A.to_scipy()
    assert abs(S - S.T).max() == 0
    r = orc.minres(A, rhs, max_iter=300, tol=1e-22)
    assert r.status == orc.OK
    assert np.allclose(S @ r.x, rhs, atol=1e-11)
    A, rhs = fx.diag_simple(8, 8)
    r = orc.minres(A, rhs, max_iter=300, tol=1e-20)
    assert r.status == orc.OK and np.allclose(r.x, 0.5, atol=1e-13)


def test_complex_fixtures_known_solution(orc):
    """tests/test_complex_solve.rs (Hermitian; MinRes, precond MinRes with real diag, precond
    BiCGStab) and tests/test_complex_solve2.rs (complex symmetric; precond BiCGStab): rhs is
    A x* with x*[i,j] = i + j i."""
    A, rhs, d, xs = fx.hermitian_grid(8, 8)
    S = A.to_scipy()
    assert abs(S - S.getH()).max() == 0
    for r in (
        orc.minres(A, rhs, max_iter=300, tol=1e-22),
        orc.minres(A, rhs, max_iter=300, tol=1e-22, pc=("diag", d)),
        orc.bicgstab(A, rhs, max_iter=300, tol=1e-22, pc=("diag", -d.astype(np.complex128) * -1)),
    ):
        assert r.status == orc.OK
        assert np.abs(r.x - xs).max() < 1e-12
    A, rhs, d, xs = fx.complex_symmetric_grid(8, 8)
    S = A.to_scipy()
    assert abs(S - S.T).max() == 0 and abs(S - S.getH()).max() > 0
    r = orc.bicgstab(A, rhs, max_iter=300, tol=1e-22, pc=("diag", d))
    assert r.status == orc.OK and np.abs(r.x - xs).max() < 1e-12


def test_csminres_unpinned_but_solves(orc):
    """CSMinRes is never run by a reference test (parity UNPINNED by the reference).  The
    complex-symmetric fixture of tests/test_complex_solve2.rs with its known x* is used instead;
    the estimate res_norm must agree with the true residual."""
    A, rhs, _, xs = fx.complex_symmetric_grid(8, 8)
    r = orc.csminres(A, rhs, max_iter=300, tol=1e-12)
    assert r.status == orc.OK
    assert np.abs(r.x - xs).max() < 1e-9
    true_rel = np.linalg.norm(A.to_scipy() @ r.x - rhs) / np.linalg.norm(rhs)
    assert true_rel < 5e-12
    # on a real symmetric matrix CSMinRes and MinRes run the same recurrence
    A2, rhs2 = fx.sym_laplacian_2d(8, 8)
    a = orc.csminres(A2, rhs2, max_iter=300, tol=1e-12)
    b = orc.minres(A2, rhs2, max_iter=300, tol=1e-12)
    assert a.iters == b.iters and np.array_equal(a.x, b.x)


# ---------------------------------------------------------------- semantics
def test_semantics_cheatsheet(orc):
    A, rhs = orc.gen_dirichlet2d(12)
    # zero rhs shortcut: x := 0, Ok((0, ||b||))  (bicg_stab.rs:55-60)
    for f in (orc.bicgstab, orc.minres, orc.csminres):
        r = f(A, np.zeros(A.n), x0=np.ones(A.n), max_iter=10, tol=1e-8)
        assert r.status == orc.OK and r.iters == 0 and np.all(r.x == 0)
    # dimension mismatch -> IncompatibleMatrixFormat (bicg_stab.rs:44-53)
    r = orc.bicgstab(A, rhs[:-1], x0=np.zeros(A.n - 1), size=A.n)
    assert r.status == orc.INCOMPATIBLE_FORMAT
    # insufficient iterations: a solve that would converge exactly at the last permitted
    # iteration still fails, the check is at the top of the next one (bicg_stab.rs:122-126)
    full = orc.bicgstab(A, rhs, max_iter=1000, tol=1e-8)
    assert full.status == orc.OK
    r = orc.bicgstab(A, rhs, max_iter=full.iters, tol=1e-8)
    assert r.status == orc.INSUFFICIENT_ITER and r.iters == full.iters
    r = orc.bicgstab(A, rhs, max_iter=full.iters + 1, tol=1e-8)
    assert r.status == orc.OK and r.iters == full.iters
    assert len(full.hist) == full.iters + 1 and full.hist[-1] == full.resid
    # MINRES its is 0-based (minres.rs:90,166)
    As, rs = fx.diag_simple(2, 2)
    r = orc.minres(As, rs, max_iter=50, tol=1e-12)
    assert r.status == orc.OK and r.iters + 1 == len(r.hist)
    # Gauss-Seidel: zero diagonal, non-square, non-CSR, max_iter == 0 (gauss_seidel.rs:16-26,52-54,72-78)
    bad = fx.kat_csr()
    r = orc.gauss_seidel(bad, np.ones(5), max_iter=5)
    assert r.status == orc.ZERO_DIAGONAL and r.iters == 0
    r = orc.gauss_seidel(A, rhs, max_iter=0)
    assert r.status == orc.INSUFFICIENT_ITER
    r = orc.gauss_seidel(A, rhs, max_iter=5, is_csr=False)
    assert r.status == orc.INCOMPATIBLE_FORMAT
    # invalid (negative definite) preconditioner for MINRES (minres.rs:236-244)
    As, rs = fx.sym_laplacian_2d(6, 6)
    r = orc.minres(As, rs, max_iter=50, tol=1e-10, pc=("diag", As.diagonal()))  # diag = -4 < 0
    assert r.status == orc.INVALID_PRECOND


def test_gs_preconditioner_definition(orc):
    """The GS/SGS preconditioner operators are project-defined (no reference analogue): forward
    sweep from zero == (D+L)^-1 r; symmetric == (D+U)^-1 D (D+L)^-1 r."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    A = orc.gen_lap3d7(6, 5, 4, shift=0.05)
    S = A.to_scipy().tocsr()
    r = np.cos(np.arange(A.n) * 0.37)
    DL = sp.tril(S, 0).tocsr()
    DU = sp.triu(S, 0).tocsr()
    D = sp.diags(S.diagonal())
    z1 = orc.gs_apply(A, r, symmetric=False)
    assert np.allclose(z1, spl.spsolve_triangular(DL, r, lower=True), rtol=1e-13, atol=1e-14)
    z2 = orc.gs_apply(A, r, symmetric=True)
    ref = spl.spsolve_triangular(DU, D @ spl.spsolve_triangular(DL, r, lower=True), lower=False)
    assert np.allclose(z2, ref, rtol=1e-12, atol=1e-13)
    # SGS-preconditioned MINRES on the shifted (indefinite) Laplacian converges (config C3, small)
    b = S @ np.ones(A.n)
    out = orc.minres(A, b, max_iter=200, tol=1e-8, pc=("gs_sym",))
    assert out.status == orc.OK and np.allclose(out.x, 1.0, atol=1e-5)


def test_generators(orc):
    A, _ = orc.gen_dirichlet2d(512)
    assert A.n == 262144 and A.nnz == 1302544  # C1
    A = orc.gen_lap3d7(16)
    S = A.to_scipy()
    assert A.nnz == 7 * 16**3 - 6 * 16**2 and abs(S - S.T).max() == 0
    assert np.all(np.diff(A.indices.astype(np.int64))[np.diff(np.repeat(np.arange(A.n), np.diff(A.indptr))) == 0] > 0)
    A = orc.gen_convdiff27(9, 8, 7)
    assert A.nnz == (3 * 9 - 2) * (3 * 8 - 2) * (3 * 7 - 2)
    S = A.to_scipy()
    assert abs(S - S.T).max() > 0
    # interior row: centre 27.75, x-1/y-1/z-1 neighbours -2/-1.5/-1.25, 27 entries, row sum 0
    row = 3 * 72 + 3 * 9 + 4
    st, en = A.indptr[row], A.indptr[row + 1]
    assert en - st == 27 and A.data[st:en].sum() == pytest.approx(0.0, abs=1e-13)
    vals = dict(zip(A.indices[st:en].tolist(), A.data[st:en].tolist()))
    assert vals[row] == 27.75 and vals[row - 1] == -2.0 and vals[row - 9] == -1.5 and vals[row - 72] == -1.25
    # partitioned generation concatenates to the full matrix
    B0, B1 = orc.gen_convdiff27(9, 8, 7, row_end=200), orc.gen_convdiff27(9, 8, 7, row_begin=200)
    assert np.array_equal(np.concatenate([B0.indices, B1.indices]), A.indices)
    assert np.array_equal(np.concatenate([B0.data, B1.data]), A.data)
