"""f32 / Complex32 through the C ABI (dtype codes SPB_F32 / SPB_C64).

The reference is generic over the four cauchy::Scalar types and dispatches s / d / c / z
(src/mkl_mat.rs:68-71,198-201,220, src/vecalg.rs:192-195); its own f32 / c32 known-answer tests are
src/vecalg.rs:647-677 and the doctests :36-46, :122-132.  Bars, as for f64: element-wise work (SpMV,
Jacobi, Gauss-Seidel sweeps, axpy-type updates) BIT-EXACT against the oracle's float restatement; every
solver's residual history bit for bit against the oracle's exact-dot flavour in float (all T::Real
quantities -- norms, Givens scalars, thresholds with f32::EPSILON -- are computed in float on both sides).
"""
import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import sprsolve_b200 as s

    s.default_context()
    return s


@pytest.fixture()
def exact(orc):
    orc.set_mode(3)
    yield orc
    orc.set_mode(0)


def single(orc, A):
    dt = np.complex64 if np.iscomplexobj(A.data) else np.float32
    return orc.Csr(A.n, A.indptr, A.indices, A.data.astype(dt), ncols=A.ncols)


def to_gpu(sp, A):
    G = sp.GpuCsrMat.new(A.indptr, A.indices, A.data, shape=(A.n, A.ncols))
    assert G.dtype == A.dtype.type
    return G


def rand_vec(n, dtype, seed=12345):
    rng = np.random.default_rng(seed)
    v = rng.uniform(-1, 1, n)
    if np.dtype(dtype).kind == "c":
        v = v + 1j * rng.uniform(-1, 1, n)
    return v.astype(dtype)


def test_vecalg_f32_c32_kats(sp):
    """src/vecalg.rs:647-677 and the doctests, through spb_vec_*."""
    va = sp.vecalg
    assert va.conj_dot(np.full(100, 2, np.float32), np.full(100, 3, np.float32)) == 600.0  # :647-650
    a, b = np.full(100, 2 + 3j, np.complex64), np.full(100, 2 - 3j, np.complex64)
    t = np.complex64(np.conj(a[0]) * b[0]) * np.float32(100)
    r = va.conj_dot(a, b)  # :652-658
    assert abs(r.real - t.real) < 1e-3 and abs(r.imag - t.imag) < 1e-3
    assert va.dot(a, b) == complex(1300.0, 0.0)  # (2+3i)(2-3i) x 100, no conjugate
    s = np.complex64(1.2 + 4.8j)  # :669-677
    v = a[0] * s
    va.scale(s, a)
    assert a.dtype == np.complex64 and np.all(np.abs(a - v) < 1e-5)
    x, y = np.full(128, 1, np.float32), np.full(128, 2, np.float32)
    va.axpby(2.0, x, -1.0, y)  # doctest :122-132
    assert np.all(y == 0)
    y = np.full(100, 0, np.float32)
    for _ in range(4):  # axpy f32 repeated 4x -> 8
        va.axpy(2.0, x[:100], y)
    assert np.all(y == 8)
    assert va.norm2(np.ones(100, np.float32)) == 10.0
    z = np.full(64, 3 - 4j, np.complex64)
    va.rscale(0.5, z)
    assert np.all(z == np.complex64(1.5 - 2j))
    out = np.zeros(64, np.complex64)
    va.conj(z, out)
    assert np.all(out == np.complex64(1.5 + 2j))


@pytest.mark.parametrize("make", [
    lambda o: o.gen_lap3d7(24, 20, 17, shift=0.05),
    lambda o: o.gen_lap3d7(15, 14, 13, shift=0.5 + 0.5j, dtype=np.complex128),
    lambda o: o.gen_convdiff27(13, 11, 10),
    lambda o: o.gen_dirichlet2d(40)[0],
    lambda o: fx.kat_csr(),  # src/mat.rs:232-255 incl. the empty row
])
def test_spmv_jacobi_gs_bit_exact(sp, orc, make, monkeypatch):
    A = single(orc, make(orc))
    x = rand_vec(A.n, A.dtype)
    yo = orc.spmv(A, x)
    for dict_knob in ("0", "1"):  # plain column stream and the column-offset dictionary
        monkeypatch.setenv("SPB_SPMV_DICT", dict_knob)
        G = to_gpu(sp, A)
        y = np.zeros(A.n, A.dtype)
        G.mul_vec(x, y)
        assert np.array_equal(y, yo), "f32 SpMV differs from the sequential float fold"
        y2 = np.zeros(A.n, A.dtype)
        d = G.mul_vec_dot(x, y2)
        assert np.array_equal(y2, yo)
        ref = np.vdot(x.astype(np.complex128), yo.astype(np.complex128))
        scale = float(np.sum(np.abs(x.astype(np.complex128)) * np.abs(yo.astype(np.complex128))))
        assert abs(d - (ref if np.iscomplexobj(x) else ref.real)) <= 1e-6 * max(scale, 1e-30)  # one float rounding of the exact sum
    monkeypatch.delenv("SPB_SPMV_DICT")
    G = to_gpu(sp, A)
    if A.n > 5:
        d = A.diagonal()
        P = sp.DiagPrecond.new(d, dtype=A.dtype)
        z = np.zeros(A.n, A.dtype)
        P.mul_vec(x, z)
        assert np.array_equal(z, orc.diag_apply(d, x))
        if np.iscomplexobj(x):  # DiagPrecond<Complex32, f32>
            dr = (np.abs(d.real) + 1).astype(np.float32)
            Pr = sp.DiagPrecond.new(dr, dtype=A.dtype)
            Pr.mul_vec(x, z)
            assert np.array_equal(z, orc.diag_apply(dr, x))
        for sym in (False, True):
            M = sp.GaussSeidelPrecond(G, symmetric=sym)
            M.mul_vec(x, z)
            assert np.array_equal(z, orc.gs_apply(A, x, sym)), "f32 Gauss-Seidel sweep differs from the sequential sweep"
            assert M.schedule_info()["poll_timeout"] == 0


def _gpu_solve(sp, A, rhs, solver, tol, max_iter, pc):
    G = to_gpu(sp, A)
    cls = {"bicgstab": sp.BiCGStab, "minres": sp.MinRes, "csminres": sp.CSMinRes}[solver]
    S = cls(G, A.n).record_history(max_iter + 1)
    x = np.zeros(A.n, dtype=A.dtype)
    M = None
    if pc is not None:
        M = {"diag": lambda: sp.DiagPrecond.new(pc[1], dtype=A.dtype), "gs_fwd": lambda: sp.GaussSeidelPrecond(G, False),
             "gs_sym": lambda: sp.GaussSeidelPrecond(G, True)}[pc[0]]()
    status, it, res = 0, None, None
    try:
        it, res = S.precond_solve(M, rhs, x, max_iter, tol) if M is not None else S.solve(rhs, x, max_iter, tol)
    except sp.InsufficientIterNum as e:
        status, it = 3, e.max_iter
    except sp.BreakDown as e:
        status, it = 4, e.its
    return status, it, res, x, S.history.copy()


CASES = [
    ("c1_64_jacobi", lambda o: o.gen_dirichlet2d(64), "bicgstab", 1e-5, 3000, "diag"),
    ("c1_32_plain", lambda o: o.gen_dirichlet2d(32), "bicgstab", 1e-4, 2000, None),
    ("c5_20_jacobi", lambda o: (lambda A: (A, o.spmv(A, np.ones(A.n))))(o.gen_convdiff27(20, 18, 16)), "bicgstab", 1e-5, 500, "diag"),
    ("c5_12_gsfwd", lambda o: (lambda A: (A, o.spmv(A, np.ones(A.n))))(o.gen_convdiff27(12, 11, 10)), "bicgstab", 1e-5, 300, "gs_fwd"),
    ("c3_16_sgs", lambda o: (lambda A: (A, o.spmv(A, np.ones(A.n))))(o.gen_lap3d7(16, shift=0.05)), "minres", 1e-4, 600, "gs_sym"),
    ("c3_20_plain", lambda o: (lambda A: (A, o.spmv(A, np.ones(A.n))))(o.gen_lap3d7(20, 18, 15, shift=0.05)), "minres", 1e-4, 800, None),
    ("c3_12_jacobi", lambda o: (lambda A: (A, o.spmv(A, np.ones(A.n))))(o.gen_lap3d7(12, shift=0.05)), "minres", 1e-4, 600, "diag"),
    ("c4_16", lambda o: (lambda A: (A, o.spmv(A, np.full(A.n, 1 + 1j))))(o.gen_lap3d7(16, shift=0.5 + 0.5j, dtype=np.complex128)), "csminres", 1e-4, 800, None),
    ("c4_10_bicg_complex", lambda o: (lambda A: (A, o.spmv(A, np.full(A.n, 1 + 1j))))(o.gen_lap3d7(10, shift=0.5 + 0.5j, dtype=np.complex128)), "bicgstab", 1e-4, 400, "diag"),
    ("hermitian_minres_realdiag", lambda o: fx.hermitian_grid(8, 8)[:3], "minres", 1e-4, 300, "extra"),
    ("csym_csminres", lambda o: fx.complex_symmetric_grid(8, 8)[:2], "csminres", 1e-4, 300, None),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_solver_history_bit_for_bit_f32(sp, exact, case, monkeypatch):
    """Same statement as tests/test_gpu_exact.py, in single precision."""
    _, build, solver, tol, max_iter, pck = case
    built = build(exact)
    A64, rhs64 = built[0], built[1]
    A = single(exact, A64)
    rhs = rhs64.astype(A.dtype)
    pc = None
    if pck == "diag":
        pc = ("diag", A.diagonal())
    elif pck == "extra":
        pc = ("diag", built[2].astype(np.float32))  # real diagonal for a complex system
    elif pck:
        pc = (pck,)
    kw = {} if solver == "csminres" else {"pc": pc}
    o = getattr(exact, solver)(A, rhs, max_iter=max_iter, tol=tol, hist_cap=max_iter + 1, **kw)
    assert o.x.dtype == A.dtype
    for fused in ("0", "1") if solver == "bicgstab" else ("",):
        if fused:
            monkeypatch.setenv("SPB_FUSED", fused)
        status, it, res, x, hist = _gpu_solve(sp, A, rhs, solver, tol, max_iter, pc)
        assert (status, it) == (o.status, o.iters)
        bad = np.flatnonzero(hist != o.hist)
        assert len(hist) == len(o.hist) and bad.size == 0, f"history differs first at iteration {bad[:1]}"
        if status == 0:
            assert res == o.resid
        assert np.array_equal(x, o.x)


def test_gauss_seidel_solver_f32(sp, exact):
    A64, rhs64 = exact.gen_dirichlet2d(12)
    A, rhs = single(exact, A64), rhs64.astype(np.float32)
    S = sp.GaussSeidel(to_gpu(sp, A)).record_history(300)
    x = np.zeros(A.n, np.float32)
    it, res = S.solve(rhs, x, 300, 1e-4)
    o = exact.gauss_seidel(A, rhs, max_iter=300, eps=1e-4, hist_cap=300)
    assert o.status == 0 and (it, res) == (o.iters, o.resid)
    assert np.array_equal(S.history, o.hist) and np.array_equal(x, o.x)
    ii, jj = np.meshgrid(np.arange(12), np.arange(12), indexing="ij")
    assert np.allclose(x, (ii + jj).ravel(), atol=5e-2)


def test_f32_semantics(sp, orc):
    """Zero rhs, f32::EPSILON zero-diagonal test (|d|^2 < 1.19e-7, src/gauss_seidel.rs:76), dtype checks."""
    A = single(orc, orc.gen_dirichlet2d(10)[0])
    G = to_gpu(sp, A)
    x = np.ones(A.n, np.float32)
    assert sp.BiCGStab(G, A.n).solve(np.zeros(A.n, np.float32), x, 10, 1e-5) == (0, 0.0) and np.all(x == 0)
    d = A.data.copy()
    d[A.indices == np.repeat(np.arange(A.n), np.diff(A.indptr))] = np.float32(2e-4)  # |d|^2 = 4e-8 < f32 eps, but >> f64 eps
    with pytest.raises(sp.ZeorDiagonalElem):
        sp.GaussSeidelPrecond(sp.GpuCsrMat.new(A.indptr, A.indices, d))
    sp.GaussSeidelPrecond(sp.GpuCsrMat.new(A.indptr, A.indices, d.astype(np.float64)))  # fine in f64
    with pytest.raises(TypeError):
        G.mul_vec(np.ones(A.n, np.float32), np.zeros(A.n, np.float64))  # output must be the matrix' scalar type


def test_f32_spmv_streams_less(sp, orc):
    """27-point f32: the dictionary kernel streams 4 bytes per non-zero (plain CSR f32: 8, f64: 12) and one 32-bit
    word per row (pattern id + low 16 bits of the row pointer) instead of the row-pointer stream."""
    A = single(orc, orc.gen_convdiff27(16, 15, 14))
    G = to_gpu(sp, A)
    info = G.plan_info()
    assert info["dictionary"] == 1
    assert info["stream_bytes"] == A.nnz * 4 + (A.n + 1) * 4 + 2 * A.n * 4
