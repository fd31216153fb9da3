"""Deterministic parity: the GPU residual history equals the exact-dot oracle BIT FOR BIT.

Every solver of the path is "element-wise work + a few long sums per iteration".  The element-wise
work of the CUDA kernels follows the reference operation for operation (bit-exact, see
test_gpu_parity.py); the long sums are carried in double-double and rounded once
(csrc/reduce.cuh), i.e. they are the exact sums correctly rounded.  The oracle's exact-dot flavour
(oracle/sprs_oracle.cpp, mode 3: same algorithms, every dot / norm = exact sum of exact products
from a Kulisch superaccumulator, rounded once) is therefore a SECOND, independent implementation
of the same mathematical object, and the two must agree in every bit of every iteration: residual
history, iteration count, returned residual and the solution vector.

  * live: the oracle runs here, on sizes it finishes in seconds -- every solver / preconditioner
    combination, including the ones whose histories are chaotic under a mere re-ordering of the
    sums (unpreconditioned BiCGStab on the Dirichlet fixture, the rho-restart path);
  * full size: BASELINE configs C1 512^2, C3 128^3 (SGS-MINRES), C4 200^3 (CSMINRES), C5 at 192^3,
    against tests/golden/exact_v1.npz (written by tests/golden/make_exact.py from the same oracle;
    minutes of CPU time, so not re-run here).  SPB_EXACT_LIVE=1 re-runs the oracle instead.

The distance between this exact-dot history and the reference's sequential-fold history is the
reference's own rounding noise; it is tabulated in profiles/r02_oracle_noise.md.
"""
import hashlib
import os

import numpy as np
import pytest

import fixtures as fx

pytestmark = pytest.mark.gpu

GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "exact_v1.npz")


@pytest.fixture(scope="module")
def sp():
    import sprsolve_b200 as s

    s.default_context()
    return s


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD_PATH)


@pytest.fixture()
def exact(orc):
    orc.set_mode(3)
    yield orc
    orc.set_mode(0)


def _sha(x):
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


def _gpu_solve(sp, A, rhs, solver, tol, max_iter, pc=None, x0=None):
    G = sp.GpuCsrMat.new(A.indptr, A.indices, A.data, shape=(A.n, A.ncols))
    cls = {"bicgstab": sp.BiCGStab, "minres": sp.MinRes, "csminres": sp.CSMinRes}[solver]
    S = cls(G, A.n).record_history(max_iter + 1)
    x = np.zeros(A.n, dtype=A.dtype) if x0 is None else np.array(x0, dtype=A.dtype)
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


def _same(gpu, o):
    status, it, res, x, hist = gpu
    assert status == o.status, (status, o.status)
    assert it == o.iters, (it, o.iters)
    assert len(hist) == len(o.hist)
    bad = np.flatnonzero(hist != o.hist)
    assert bad.size == 0, f"history differs first at iteration {bad[0]}: {hist[bad[0]]!r} vs {o.hist[bad[0]]!r}"
    if status == 0:
        assert res == o.resid
    assert np.array_equal(x, o.x), "solution vectors differ"


LIVE = [
    # (id, matrix builder -> (A, rhs, extras), solver, tol, max_iter, pc builder)
    ("c1_256_jacobi", lambda o: o.gen_dirichlet2d(256), "bicgstab", 1e-8, 10000, lambda A, e: ("diag", A.diagonal())),
    ("c1_96_plain", lambda o: o.gen_dirichlet2d(96), "bicgstab", 1e-8, 5000, None),  # chaotic under re-ordering
    ("c5_40_jacobi", lambda o: _ones_rhs(o, o.gen_convdiff27(40, 36, 33)), "bicgstab", 1e-8, 1000, lambda A, e: ("diag", A.diagonal())),
    ("c5_20_gsfwd", lambda o: _ones_rhs(o, o.gen_convdiff27(20, 18, 16)), "bicgstab", 1e-8, 500, lambda A, e: ("gs_fwd",)),
    ("c5_20_plain", lambda o: _ones_rhs(o, o.gen_convdiff27(20, 18, 16)), "bicgstab", 1e-8, 500, None),
    ("restart", lambda o: _ones_rhs(o, o.gen_lap3d7(6, shift=0.0)), "bicgstab", 1e-30, 400, None),  # rho restart, src/bicg_stab.rs:304-318
    ("c3_32_sgs", lambda o: _ones_rhs(o, o.gen_lap3d7(32, shift=0.05)), "minres", 1e-8, 1000, lambda A, e: ("gs_sym",)),
    ("c3_40_plain", lambda o: _ones_rhs(o, o.gen_lap3d7(40, 37, 31, shift=0.05)), "minres", 1e-8, 2000, None),
    ("c3_24_jacobi", lambda o: _ones_rhs(o, o.gen_lap3d7(24, shift=0.05)), "minres", 1e-8, 1000, lambda A, e: ("diag", A.diagonal())),
    ("c4_40", lambda o: _ones_rhs(o, o.gen_lap3d7(40, shift=0.5 + 0.5j, dtype=np.complex128), 1 + 1j), "csminres", 1e-8, 2000, None),
    ("hermitian_minres", lambda o: fx.hermitian_grid(8, 8)[:3], "minres", 1e-10, 300, None),
    ("hermitian_minres_realdiag", lambda o: fx.hermitian_grid(8, 8)[:3], "minres", 1e-10, 300, lambda A, e: ("diag", e)),
    ("hermitian_bicg_realdiag", lambda o: fx.hermitian_grid(8, 8)[:3], "bicgstab", 1e-12, 300, lambda A, e: ("diag", e)),
    ("csym_bicg_cdiag", lambda o: fx.complex_symmetric_grid(8, 8)[:3], "bicgstab", 1e-12, 300, lambda A, e: ("diag", e)),
    ("csym_csminres", lambda o: fx.complex_symmetric_grid(8, 8)[:3], "csminres", 1e-10, 300, None),
    ("sym2d_minres", lambda o: fx.sym_laplacian_2d(8, 8), "minres", 1e-10, 300, None),
    ("c4_16_bicg_complex", lambda o: _ones_rhs(o, o.gen_lap3d7(16, shift=0.5 + 0.5j, dtype=np.complex128), 1 + 1j), "bicgstab", 1e-8, 600, None),
]


def _ones_rhs(o, A, v=1.0):
    return A, o.spmv(A, np.full(A.n, v, dtype=A.dtype))


@pytest.mark.parametrize("case", LIVE, ids=[c[0] for c in LIVE])
def test_history_bit_for_bit_live(sp, exact, case, monkeypatch):
    _, build, solver, tol, max_iter, pcb = case
    built = build(exact)
    A, rhs = built[0], built[1]
    extra = built[2] if len(built) > 2 else None
    pc = pcb(A, extra) if pcb else None
    kw = {} if solver == "csminres" else {"pc": pc}
    o = getattr(exact, solver)(A, rhs, max_iter=max_iter, tol=tol, hist_cap=max_iter + 1, **kw)
    # BiCGStab has two device paths: the multi-kernel loop and the single cooperative kernel that takes
    # L2-resident systems (csrc/bicgstab.cu) -- both must reproduce the oracle
    for fused in ("0", "1") if solver == "bicgstab" else ("",):
        if fused:
            monkeypatch.setenv("SPB_FUSED", fused)
        _same(_gpu_solve(sp, A, rhs, solver, tol, max_iter, pc), o)


FUSED_MODES = [  # (id, environment) -- csrc/bicgstab.cu: bicg_fused_kernel<.., BLOCK, MODE>
    ("grid_global_256", {"SPB_FUSED_SMEM": "0", "SPB_FUSED_CLUSTER": "0", "SPB_FUSED_BLOCK": "256"}),
    ("grid_global_512", {"SPB_FUSED_SMEM": "0", "SPB_FUSED_CLUSTER": "0", "SPB_FUSED_BLOCK": "512"}),
    ("grid_smem_256", {"SPB_FUSED_CLUSTER": "0", "SPB_FUSED_BLOCK": "256"}),
    ("grid_smem_512", {"SPB_FUSED_CLUSTER": "0", "SPB_FUSED_BLOCK": "512"}),
    ("cluster_256", {"SPB_FUSED_CLUSTER": "1", "SPB_FUSED_BLOCK": "256"}),
    ("cluster_512", {"SPB_FUSED_CLUSTER": "1", "SPB_FUSED_BLOCK": "512"}),
]


@pytest.mark.parametrize("mode", FUSED_MODES, ids=[m[0] for m in FUSED_MODES])
def test_fused_kernel_modes_bit_for_bit(sp, exact, mode, monkeypatch):
    """Every mode of the single-kernel BiCGStab -- cooperative grid with global or shared-memory vectors, and the
    one-cluster mode (distributed shared memory, matrix slice + window in shared memory) -- reproduces the exact-dot
    oracle in every bit: Jacobi on the reference Dirichlet matrix (several CTAs, 1-2 rows per thread), a 27-point
    matrix whose window is wider than a CTA's rows, the rho-restart path across several CTAs, a complex system without
    preconditioner, and a system of one partial CTA."""
    monkeypatch.setenv("SPB_FUSED", "1")
    for k, v in mode[1].items():
        monkeypatch.setenv(k, v)
    cases = [
        (exact.gen_dirichlet2d(100), "diag", 1e-8, 2000),
        (_ones_rhs(exact, exact.gen_convdiff27(14, 13, 12)), "diag", 1e-8, 500),
        (_ones_rhs(exact, exact.gen_lap3d7(12, shift=0.0)), None, 1e-30, 300),  # rho restart, 1728 rows
        (_ones_rhs(exact, exact.gen_lap3d7(11, 10, 9, shift=0.5 + 0.5j, dtype=np.complex128), 1 + 1j), None, 1e-8, 600),
        (_ones_rhs(exact, exact.gen_lap3d7(5, 4, 3, shift=0.05)), "diag", 1e-10, 200),
    ]
    for (A, rhs), pck, tol, max_iter in cases:
        pc = ("diag", A.diagonal()) if pck else None
        o = exact.bicgstab(A, rhs, max_iter=max_iter, tol=tol, hist_cap=max_iter + 1, pc=pc)
        _same(_gpu_solve(sp, A, rhs, "bicgstab", tol, max_iter, pc), o)


def test_gauss_seidel_solver_bit_for_bit(sp, exact):
    """GaussSeidel::solve (src/gauss_seidel.rs:33-140): the sweeps are bit-exact and the per-sweep
    residual norm / the b-norm are exactly rounded on both sides."""
    for A, rhs, its, eps in ((*exact.gen_dirichlet2d(12), 2000, 1e-9), (*exact.gen_dirichlet2d(24), 300, 1e-9),
                             (*_ones_rhs(exact, exact.gen_convdiff27(10, 9, 8)), 200, 1e-10)):
        G = sp.GpuCsrMat.new(A.indptr, A.indices, A.data)
        S = sp.GaussSeidel(G).record_history(its)
        x = np.zeros(A.n)
        o = exact.gauss_seidel(A, rhs, max_iter=its, eps=eps, hist_cap=its)
        try:
            it, res = S.solve(rhs, x, its, eps)
            assert o.status == 0 and (it, res) == (o.iters, o.resid)
        except sp.InsufficientIterNum as e:
            assert o.status == 3 and e.max_iter == its
        assert np.array_equal(S.history, o.hist) and np.array_equal(x, o.x)


def test_vecalg_reductions_are_exactly_rounded(sp, exact):
    """dot / conj_dot / norm2 through the ABI == the exact sum rounded once, for ill-conditioned sums too."""
    rng = np.random.default_rng(2024)
    for n in (1, 7, 1000, 262144 + 3):
        sc = 10.0 ** rng.integers(-12, 12, size=n)
        x, y = rng.standard_normal(n) * sc, rng.standard_normal(n) / sc
        y[: n // 2] *= -1.0
        assert sp.vecalg.conj_dot(x, y) == exact.conj_dot(x, y)
        assert sp.vecalg.dot(x, y) == exact.dot(x, y)
        assert sp.vecalg.norm2(x) == exact.norm2(x)
        xc, yc = x + 1j * y[::-1], y - 0.5j * x
        assert sp.vecalg.conj_dot(xc, yc) == exact.conj_dot(xc, yc)
        assert sp.vecalg.dot(xc, yc) == exact.dot(xc, yc)
        assert sp.vecalg.norm2(xc) == exact.norm2(xc)


# ------------------------------------------------------------------ full BASELINE sizes
def _dev_solve(sp, G, solver, M, ones_value, max_iter, tol):
    """rhs = A * ones on the device, x0 = 0, device-resident solve; returns (status, it, res, x host, hist)."""
    import torch

    n = G.n_local
    cplx = np.dtype(G.dtype).kind == "c"
    tdt = torch.complex128 if cplx else torch.float64
    dev = torch.device("cuda:0")
    ones = torch.full((n,), ones_value, dtype=tdt, device=dev)
    rhs = torch.empty(n, dtype=tdt, device=dev)
    x = torch.zeros(n, dtype=tdt, device=dev)
    torch.cuda.synchronize()
    G.mul_vec_dev(ones.data_ptr(), rhs.data_ptr())
    cls = {"bicgstab": sp.BiCGStab, "minres": sp.MinRes, "csminres": sp.CSMinRes}[solver]
    S = cls(G, n).record_history(max_iter + 1)
    it, res = S.solve_dev(rhs.data_ptr(), x.data_ptr(), max_iter, tol, precond=M)
    G.ctx.synchronize()
    return it, res, x.cpu().numpy(), S.history.copy()


def _check_gold(gold, name, it, res, x, hist):
    g = lambda k: gold[f"{name}.exact.{k}"]  # noqa: E731
    assert int(g("status")) == 0
    assert it == int(g("iters")), (it, int(g("iters")))
    gh = g("hist")
    assert len(hist) == len(gh)
    bad = np.flatnonzero(hist != gh)
    assert bad.size == 0, f"{name}: history differs first at iteration {bad[0]} of {len(gh)}"
    assert res == float(g("resid"))
    assert np.array_equal(x[:16], g("x_head"))
    assert _sha(x) == str(g("x_sha256")), f"{name}: solution differs from the exact-dot oracle's"


def _live(exact, name):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_exact", os.path.join(os.path.dirname(GOLD_PATH), "make_exact.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    A, rhs, solver, kw = m.CASES[name][0]()
    exact.set_mode(3)
    return getattr(exact, solver)(A, rhs, hist_cap=kw["max_iter"] + 1, **kw)


def _gold_or_live(gold, exact, name, it, res, x, hist):
    if os.environ.get("SPB_EXACT_LIVE") == "1":
        o = _live(exact, name)
        assert (o.status, o.iters, o.resid) == (0, it, res) and np.array_equal(o.hist, hist) and np.array_equal(o.x, x)
    _check_gold(gold, name, it, res, x, hist)


def test_c1_full_size_bit_for_bit(sp, gold, exact, monkeypatch):
    """Config C1 at full size: 512^2 reference Dirichlet matrix (src/main.rs:53-88), Jacobi, rtol 1e-8.
    1035 iterations -- the sequential-fold reference takes 933, its OpenMP re-orderings 842..1108:
    that spread is rounding noise (profiles/r02_oracle_noise.md); this comparison has none."""
    A, rhs = exact.gen_dirichlet2d(512)
    for fused in ("1", "0"):  # single-kernel solve (the default at this size) and the multi-kernel loop
        monkeypatch.setenv("SPB_FUSED", fused)
        status, it, res, x, hist = _gpu_solve(sp, A, rhs, "bicgstab", 1e-8, 10000, ("diag", A.diagonal()))
        assert status == 0
        _gold_or_live(gold, exact, "c1_512", it, res, x, hist)


def test_c3_full_size_bit_for_bit(sp, gold, exact):
    """Config C3 at full size: shifted 7-point Laplacian 128^3, symmetric Gauss-Seidel preconditioner
    (block-wavefront sweep), MINRES to 1e-8 -- and the unpreconditioned solve."""
    G = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, 128, 128, 128, params=(0.05,))
    M = sp.GaussSeidelPrecond(G, symmetric=True)
    _gold_or_live(gold, exact, "c3_128", *_dev_solve(sp, G, "minres", M, 1.0, 10000, 1e-8))
    _gold_or_live(gold, exact, "c3_128_plain", *_dev_solve(sp, G, "minres", None, 1.0, 10000, 1e-8))


def test_c4_full_size_bit_for_bit(sp, gold, exact):
    """Config C4 at full size: complex-symmetric Helmholtz 200^3, CSMinRes to 1e-8."""
    G = sp.GpuCsrMat.from_stencil(sp.STENCIL_LAP3D7, 200, 200, 200, params=(0.5, 0.5), dtype=np.complex128)
    _gold_or_live(gold, exact, "c4_200", *_dev_solve(sp, G, "csminres", None, 1 + 1j, 10000, 1e-8))


@pytest.mark.parametrize("g", [48, 96, 192])
def test_c5_bit_for_bit(sp, gold, exact, g):
    """Config C5 (27-point convection-diffusion, Jacobi-BiCGStab to 1e-8) at 48^3 / 96^3 / 192^3 (7.1 M
    rows, 189 M non-zeros; the 512^3 matrix is beyond a CPU oracle -- 44 GB and hours)."""
    G = sp.GpuCsrMat.from_stencil(sp.STENCIL_CONVDIFF27, g, g, g, params=(1.0, 0.5, 0.25))
    assert G.plan_info()["dictionary"] == 1
    M = sp.DiagPrecond.from_matrix(G)
    _gold_or_live(gold, exact, f"c5_{g}", *_dev_solve(sp, G, "bicgstab", M, 1.0, 10000, 1e-8))
