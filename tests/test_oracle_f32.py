"""f32 / Complex32 restatement in the CPU oracle (the reference is
generic over all four cauchy::Scalar types, src/mkl_mat.rs:68-71; the GPU library implements f64 /
Complex64 so far).  CPU only.  Pinned by the reference's own f32 / c32 known-answer tests; beyond
them the f32 code paths must agree with the f64 restatement to float accuracy and must behave like
f32 arithmetic (sequential float folds, f32::EPSILON thresholds), which is what is checked here."""
import numpy as np
import pytest


def test_vecalg_f32_c32_kats(orc):
    # src/vecalg.rs:647-650: conj_dot of 2_f32 x 3_f32, len 100 == 600 exactly
    assert orc.conj_dot(np.full(100, 2, np.float32), np.full(100, 3, np.float32)) == np.float32(600)
    # :652-658: c32 conj_dot == conj(a0) * b0 * 100
    a = np.full(100, 2 + 3j, np.complex64)
    b = np.full(100, 2 - 3j, np.complex64)
    r = orc.conj_dot(a, b)
    t = np.complex64(np.conj(a[0]) * b[0]) * np.float32(100)
    assert abs(r.real - t.real) < 1e-3 and abs(r.imag - t.imag) < 1e-3
    # :669-677: scale by a c32
    a = np.full(100, 2 + 3j, np.complex64)
    s = np.complex64(1.2 + 4.8j)
    v = a[0] * s
    orc.scale(s, a)
    assert np.all(np.abs(a.real - v.real) < 1e-5) and np.all(np.abs(a.imag - v.imag) < 1e-5)
    # :694-697: dot f32
    assert orc.dot(np.full(100, 2, np.float32), np.full(100, 3, np.float32)) == np.float32(600)
    # doctests :122-132 (axpby, f32) and :36-46-style axpy
    x = np.full(128, 1, np.float32)
    y = np.full(128, 2, np.float32)
    orc.axpby(2.0, x, -1.0, y)
    assert np.all(y == 0)
    y = np.full(128, 2, np.float32)
    orc.axpy(2.0, x, y)
    assert np.all(y == 4)
    assert orc.norm2(np.ones(100, np.float32)) == pytest.approx(10.0, abs=1e-6)
    assert orc.norm2(np.full(50, 1 + 1j, np.complex64)) == pytest.approx(10.0, abs=1e-5)


def test_f32_folds_are_float_folds(orc):
    """The sums are sequential FLOAT folds (vecalg.rs:557-568, 601-605 with T = f32): they reproduce
    a numpy float32 loop bit for bit and differ from the f64 fold rounded to float."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 4000).astype(np.float32)
    y = rng.uniform(-1, 1, 4000).astype(np.float32)
    acc = np.float32(0)
    for a, b in zip(x, y):
        acc = np.float32(acc + np.float32(a * b))
    assert orc.dot(x, y) == acc == orc.conj_dot(x, y)
    sq = np.float32(0)
    for a in x:
        sq = np.float32(sq + np.float32(a * a))
    assert np.float32(orc.norm2(x)) == np.sqrt(sq)
    assert orc.dot(x, y) != np.float32(orc.dot(x.astype(np.float64), y.astype(np.float64)))  # not a rounded f64 sum


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_operators_f32_match_f64_to_float_accuracy(orc, dtype):
    wide = np.complex128 if dtype is np.complex64 else np.float64
    A64 = orc.gen_lap3d7(7, 6, 5, shift=(0.3 + 0.2j) if dtype is np.complex64 else 0.3, dtype=wide)
    A = orc.Csr(A64.n, A64.indptr, A64.indices, A64.data.astype(dtype))
    assert A.dtype == dtype
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, A.n).astype(dtype)
    if dtype is np.complex64:
        x = (x + 1j * rng.uniform(-1, 1, A.n)).astype(dtype)
    y = orc.spmv(A, x)
    assert y.dtype == dtype
    assert np.allclose(y, orc.spmv(A64, x.astype(wide)), rtol=2e-5, atol=2e-5)
    y2, d = orc.spmv_dot(A, x)
    assert np.array_equal(y2, y) and abs(d - np.vdot(x.astype(wide), y.astype(wide))) < 1e-3
    for sym in (False, True):
        z = orc.gs_apply(A, x, sym)
        assert z.dtype == dtype
        assert np.allclose(z, orc.gs_apply(A64, x.astype(wide), sym), rtol=1e-4, atol=1e-4)
    dj = orc.diag_apply(A.diagonal(), x)
    assert np.allclose(dj, orc.diag_apply(A64.diagonal(), x.astype(wide)), rtol=1e-5, atol=1e-6)


def test_solvers_f32(orc):
    """BiCGStab / MINRES / CSMinRes / GaussSeidel in f32 converge to float accuracy on the reference's
    fixtures.  (BiCGStab's test quantity is the recursively updated residual, which keeps shrinking
    below float resolution -- as in the reference -- while the true residual stalls there.)"""
    A64, rhs64 = orc.gen_dirichlet2d(20)
    A = orc.Csr(A64.n, A64.indptr, A64.indices, A64.data.astype(np.float32))
    rhs = rhs64.astype(np.float32)
    ii, jj = np.meshgrid(np.arange(20), np.arange(20), indexing="ij")
    exact = (ii + jj).ravel()
    o = orc.bicgstab(A, rhs, max_iter=1500, tol=1e-5, pc=("diag", A.diagonal()))
    assert o.status == orc.OK and o.x.dtype == np.float32
    assert np.allclose(o.x, exact, atol=5e-3)
    o64 = orc.bicgstab(A64, rhs64, max_iter=1500, tol=1e-5, pc=("diag", A64.diagonal()))
    assert abs(o.iters - o64.iters) <= max(3, o64.iters // 4)
    assert np.allclose(o.hist[:5], o64.hist[:5], rtol=1e-3)
    tight = orc.bicgstab(A, rhs, max_iter=400, tol=1e-12, pc=("diag", A.diagonal()))
    assert tight.status == orc.OK and tight.resid < 1e-12
    true_rel = np.linalg.norm(A64.to_scipy() @ tight.x.astype(np.float64) - rhs64) / np.linalg.norm(rhs64)
    assert 1e-9 < true_rel < 1e-5  # the f32 iterate cannot do better than float resolution
    # MINRES on the shifted Laplacian (symmetric), SGS-preconditioned and plain
    L64 = orc.gen_lap3d7(8, 8, 8, shift=0.05)
    L = orc.Csr(L64.n, L64.indptr, L64.indices, L64.data.astype(np.float32))
    b = orc.spmv(L, np.ones(L.n, np.float32))
    for pc in (None, ("gs_sym",)):
        m = orc.minres(L, b, max_iter=400, tol=1e-4, pc=pc)
        assert m.status == orc.OK and np.allclose(m.x, 1.0, atol=5e-2)
    # complex-symmetric Helmholtz in Complex32 with CSMinRes
    H64 = orc.gen_lap3d7(6, 6, 6, shift=0.5 + 0.5j, dtype=np.complex128)
    H = orc.Csr(H64.n, H64.indptr, H64.indices, H64.data.astype(np.complex64))
    xs = (np.cos(0.37 * np.arange(H.n)) + 1j * np.sin(0.11 * np.arange(H.n))).astype(np.complex64)
    c = orc.csminres(H, orc.spmv(H, xs), max_iter=600, tol=1e-4)
    assert c.status == orc.OK and np.allclose(c.x, xs, atol=5e-2)
    # stationary Gauss-Seidel, f32: src/gauss_seidel.rs semantics (absolute residual, zero-diagonal test with f32 eps)
    g = orc.gauss_seidel(A, rhs, max_iter=300, eps=1e-4)
    assert g.status == orc.OK and np.allclose(g.x, exact, atol=5e-2)
