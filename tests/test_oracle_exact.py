"""CPU tests of the oracle's exact-dot flavour (mode 3) -- the checker the deterministic GPU parity
tests (tests/test_gpu_exact.py) are anchored on -- and of the committed full-size goldens."""
import os
from fractions import Fraction

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_v1.npz")


def _exact(orc):
    orc.set_mode(3)
    return orc


def test_exact_dots_equal_rational_arithmetic(orc):
    """dot / conj_dot / norm2 in mode 3 are the exact sums (python Fractions) rounded once -- also for
    ill-conditioned, cancelling and wide-dynamic-range data, complex and f32."""
    rng = np.random.default_rng(7)
    try:
        _exact(orc)
        for trial in range(60):
            n = int(rng.integers(1, 300))
            sc = 10.0 ** rng.integers(-30, 30, size=n)
            x = rng.standard_normal(n) * sc
            y = rng.standard_normal(n) / sc * 10.0 ** rng.integers(-8, 8, size=n)
            if trial % 3 == 0:
                y[: n // 2] = -y[: n // 2]
            e = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(x, y))
            assert orc.conj_dot(x, y) == float(e) and orc.dot(x, y) == float(e)
            assert orc.norm2(x) == float(np.sqrt(float(sum(Fraction(float(a)) ** 2 for a in x))))
            xc = x[: n // 2 + 1] + 1j * y[: n // 2 + 1]
            yc = y[::-1][: n // 2 + 1] - 0.5j * x[::-1][: n // 2 + 1]
            F = lambda v: Fraction(float(v))  # noqa: E731
            re = sum(F(a.real) * F(b.real) + F(a.imag) * F(b.imag) for a, b in zip(xc, yc))
            im = sum(F(a.real) * F(b.imag) - F(a.imag) * F(b.real) for a, b in zip(xc, yc))
            assert orc.conj_dot(xc, yc) == complex(float(re), float(im))
            re = sum(F(a.real) * F(b.real) - F(a.imag) * F(b.imag) for a, b in zip(xc, yc))
            im = sum(F(a.real) * F(b.imag) + F(a.imag) * F(b.real) for a, b in zip(xc, yc))
            assert orc.dot(xc, yc) == complex(float(re), float(im))
            xs, ys = x.astype(np.float32), rng.standard_normal(n).astype(np.float32)
            e = sum(Fraction(float(a)) * Fraction(float(b)) for a, b in zip(xs, ys))
            assert orc.conj_dot(xs, ys) == np.float32(float(e))
    finally:
        orc.set_mode(0)


def test_level_parallel_gauss_seidel_is_the_sequential_sweep(orc):
    """Modes >= 1 run the Gauss-Seidel sweeps level by level on all cores; bit-identical to the
    sequential loop of src/gauss_seidel.rs:111-125 (stencils, random patterns, complex, f32)."""
    import scipy.sparse as sps

    rng = np.random.default_rng(3)
    M = (sps.random(1500, 1500, 0.004, random_state=5, format="csr") + sps.diags(np.full(1500, 3.0))).tocsr()
    M.sort_indices()
    mats = [orc.gen_lap3d7(17, 13, 11, shift=0.05), orc.gen_convdiff27(12, 11, 9), orc.Csr(1500, M.indptr, M.indices, M.data),
            orc.Csr(1500, M.indptr, M.indices, M.data * (1 + 0.25j)), orc.Csr(1500, M.indptr, M.indices, M.data.astype(np.float32))]
    try:
        for A in mats:
            v = rng.standard_normal(A.n).astype(A.dtype)
            for sym in (False, True):
                orc.set_mode(0)
                ref = orc.gs_apply(A, v, sym)
                for mode in (1, 3):
                    orc.set_mode(mode)
                    assert np.array_equal(orc.gs_apply(A, v, sym), ref)
    finally:
        orc.set_mode(0)


def test_exact_flavour_is_thread_count_independent_and_close_to_sequential(orc):
    A = orc.gen_convdiff27(20, 18, 16)
    rhs = orc.spmv(A, np.ones(A.n))
    pc = ("diag", A.diagonal())
    try:
        seq = orc.bicgstab(A, rhs, max_iter=500, tol=1e-8, pc=pc, hist_cap=501)
        _exact(orc)
        runs = []
        for thr in (1, 3, orc.max_threads()):
            orc.set_threads(thr)
            runs.append(orc.bicgstab(A, rhs, max_iter=500, tol=1e-8, pc=pc, hist_cap=501))
        for r in runs[1:]:
            assert r.iters == runs[0].iters and np.array_equal(r.hist, runs[0].hist) and np.array_equal(r.x, runs[0].x)
        # same algorithm, only the rounding of the sums differs: the first iterations agree to ~1e-13
        assert np.all(np.abs(runs[0].hist[:8] - seq.hist[:8]) <= 1e-12 * seq.hist[:8])
        assert abs(runs[0].iters - seq.iters) <= 3
        # MINRES with the symmetric Gauss-Seidel preconditioner and the stationary solver as well
        B = orc.gen_lap3d7(14, shift=0.05)
        rb = orc.spmv(B, np.ones(B.n))
        m = [None, None]
        for k, thr in enumerate((1, orc.max_threads())):
            orc.set_threads(thr)
            m[k] = orc.minres(B, rb, max_iter=400, tol=1e-8, pc=("gs_sym",), hist_cap=400)
        assert m[0].status == 0 and m[0].iters == m[1].iters and np.array_equal(m[0].hist, m[1].hist) and np.array_equal(m[0].x, m[1].x)
        g = orc.gauss_seidel(A, rhs, max_iter=50, eps=1e-6, hist_cap=50)
        orc.set_mode(0)
        g0 = orc.gauss_seidel(A, rhs, max_iter=50, eps=1e-6, hist_cap=50)
        assert np.array_equal(g.x[: 10], g0.x[:10]) and np.allclose(g.hist, g0.hist[: len(g.hist)], rtol=1e-12)
    finally:
        orc.set_mode(0)
        orc.set_threads(orc.max_threads())


def test_committed_exact_goldens_are_reproducible(orc):
    """tests/golden/exact_v1.npz (full-size histories) comes from tests/golden/make_exact.py: re-derive
    the cheap entries here -- the whole c5_48 case and the first 150 iterations of C1 512^2."""
    import hashlib

    gold = np.load(GOLD)
    for name in ("c1_512", "c3_128", "c3_128_plain", "c4_200", "c5_48", "c5_96", "c5_192"):
        assert int(gold[f"{name}.exact.status"]) == 0 and len(gold[f"{name}.exact.hist"]) > 10
    assert int(gold["c1_512.exact.iters"]) == 1035 and int(gold["c1_512.seq.iters"]) == 933
    try:
        _exact(orc)
        A = orc.gen_convdiff27(48)
        rhs = orc.spmv(A, np.ones(A.n), parallel=True)
        o = orc.bicgstab(A, rhs, max_iter=10000, tol=1e-8, pc=("diag", A.diagonal()), hist_cap=10001)
        assert o.iters == int(gold["c5_48.exact.iters"]) and o.resid == float(gold["c5_48.exact.resid"])
        assert np.array_equal(o.hist, gold["c5_48.exact.hist"])
        assert hashlib.sha256(o.x.tobytes()).hexdigest() == str(gold["c5_48.exact.x_sha256"])
        A, rhs = orc.gen_dirichlet2d(512)
        o = orc.bicgstab(A, rhs, max_iter=150, tol=1e-8, pc=("diag", A.diagonal()), hist_cap=151)
        assert np.array_equal(o.hist, gold["c1_512.exact.hist"][: len(o.hist)])
    finally:
        orc.set_mode(0)
