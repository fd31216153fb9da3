"""ctypes loader for the CPU ORACLE (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package ``sprsolve_b200`` never does (tests enforce that).

The oracle restates the reference's non-MKL, sequential code paths (see sprs_oracle.h for the
file:line map).  Arrays: float64 (real) or complex128 (interleaved re,im); int32 column indices;
int64 row pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK, INCOMPATIBLE_FORMAT, ZERO_DIAGONAL, INSUFFICIENT_ITER, BREAKDOWN, INVALID_PRECOND = range(6)
PC_NONE, PC_DIAG, PC_DIAG_REAL, PC_GS_FWD, PC_GS_SYM, PC_SOR_FWD, PC_SSOR = range(7)

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (g++ -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "sprs_oracle.cpp")
    hdr = os.path.join(_HERE, "sprs_oracle.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or (os.path.exists(src) and os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_i64 = C.c_int64
_dbl = C.c_double
_p = C.c_void_p


def _declare(L):
    L.orc_norm2_d.restype = _dbl
    L.orc_norm2_z.restype = _dbl
    L.orc_dot_d.restype = _dbl
    L.orc_conj_dot_d.restype = _dbl
    L.orc_conj_dot_s.restype = C.c_float
    for name in ("orc_norm2_s", "orc_norm2_c", "orc_dot_s", "orc_spmv_dot_s"):
        getattr(L, name).restype = C.c_float
    for name in (
        "orc_gen_dirichlet2d",
        "orc_gen_lap3d7_d",
        "orc_gen_lap3d7_z",
        "orc_gen_convdiff27_d",
    ):
        getattr(L, name).restype = _i64
    L.orc_max_threads.restype = C.c_int
    L.orc_get_mode.restype = C.c_int


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _is_c(a) -> bool:
    return np.iscomplexobj(a)


def _sfx(dtype) -> str:
    """d / z (f64, Complex64) and s / c (f32, Complex32 -- SURVEY.md section 8f rank 2; the GPU
    library does not implement them yet, the oracle already restates them)."""
    dt = np.dtype(dtype)
    if dt.kind == "c":
        return "c" if dt.itemsize == 8 else "z"
    return "s" if dt.itemsize == 4 else "d"


def _is_single(dtype) -> bool:
    return _sfx(dtype) in ("s", "c")


def _real_dtype(dtype):
    return np.float32 if _is_single(dtype) else np.float64


@dataclass
class Csr:
    """Raw 3-array CSR (the layout sprs::CsMatI::into_raw_storage hands to MklMat::new,
    src/mkl_mat.rs:41)."""

    n: int
    indptr: np.ndarray  # int64 [n+1]
    indices: np.ndarray  # int32 [nnz]
    data: np.ndarray  # float64 | complex128 [nnz]
    ncols: int | None = None

    def __post_init__(self):
        self.indptr = np.ascontiguousarray(self.indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.int32)
        d = np.asarray(self.data)
        if d.dtype in (np.float32, np.complex64):  # f32 / Complex32 are kept (oracle only)
            dt = d.dtype
        else:
            dt = np.complex128 if _is_c(self.data) else np.float64
        self.data = np.ascontiguousarray(self.data, dtype=dt)
        if self.ncols is None:
            self.ncols = self.n

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])

    @property
    def dtype(self):
        return self.data.dtype

    def diagonal(self) -> np.ndarray:
        d = np.zeros(self.n, dtype=self.dtype)
        rows = np.repeat(np.arange(self.n), np.diff(self.indptr))
        m = rows == self.indices
        d[rows[m]] = self.data[m]
        return d

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=(self.n, self.ncols))


def set_mode(mode: int) -> None:
    lib().orc_set_mode(int(mode))


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------- operators
def spmv(A: Csr, x: np.ndarray, parallel: bool = False) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=A.dtype)
    y = np.empty(A.n, dtype=A.dtype)
    fn = getattr(lib(), f"orc_spmv_{'par_' if parallel else ''}{_sfx(A.dtype)}")
    fn(_i64(A.n), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data), _ptr(x), _ptr(y))
    return y


def spmv_csc(nrows, ncols, indptr, indices, data, x) -> np.ndarray:
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    dt = np.complex128 if (np.iscomplexobj(data) or np.iscomplexobj(x)) else np.float64
    data = np.ascontiguousarray(data, dtype=dt)
    x = np.ascontiguousarray(x, dtype=dt)
    y = np.empty(nrows, dtype=dt)
    fn = lib().orc_spmv_csc_z if dt is np.complex128 else lib().orc_spmv_csc_d
    fn(_i64(nrows), _i64(ncols), _ptr(indptr), _ptr(indices), _ptr(data), _ptr(x), _ptr(y))
    return y


def spmv_dot(A: Csr, x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=A.dtype)
    y = np.empty(A.n, dtype=A.dtype)
    sfx = _sfx(A.dtype)
    if sfx == "s":
        r = lib().orc_spmv_dot_s(_i64(A.n), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data), _ptr(x), _ptr(y))
        return y, np.float32(r)
    out = np.zeros(2, dtype=_real_dtype(A.dtype))
    getattr(lib(), f"orc_spmv_dot_{sfx}")(
        _i64(A.n), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data), _ptr(x), _ptr(y), _ptr(out)
    )
    if sfx == "c":
        return y, np.complex64(complex(out[0], out[1]))
    return y, (complex(out[0], out[1]) if _is_c(A.data) else float(out[0]))


# --------------------------------------------------------------------------- vecalg
def _c2(a: complex) -> np.ndarray:
    return np.array([complex(a).real, complex(a).imag], dtype=np.float64)


def norm2(x) -> float:
    x = np.ascontiguousarray(x)
    return float(getattr(lib(), f"orc_norm2_{_sfx(x.dtype)}")(_i64(x.size), _ptr(x)))


def _c2f(a: complex) -> np.ndarray:
    return np.array([complex(a).real, complex(a).imag], dtype=np.float32)


def dot(x, y):
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    if _sfx(x.dtype) == "s":
        return np.float32(lib().orc_dot_s(_i64(x.size), _ptr(x), _ptr(y)))
    if _sfx(x.dtype) == "c":
        out = np.zeros(2, np.float32)
        lib().orc_dot_c(_i64(x.size), _ptr(x), _ptr(y), _ptr(out))
        return np.complex64(complex(out[0], out[1]))
    if _is_c(x):
        out = np.zeros(2)
        lib().orc_dot_z(_i64(x.size), _ptr(x), _ptr(y), _ptr(out))
        return complex(out[0], out[1])
    return float(lib().orc_dot_d(_i64(x.size), _ptr(x), _ptr(y)))


def conj_dot(x, y):
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    if _sfx(x.dtype) == "s":
        return np.float32(lib().orc_conj_dot_s(_i64(x.size), _ptr(x), _ptr(y)))
    if _sfx(x.dtype) == "c":
        out = np.zeros(2, np.float32)
        lib().orc_conj_dot_c(_i64(x.size), _ptr(x), _ptr(y), _ptr(out))
        return np.complex64(complex(out[0], out[1]))
    if _is_c(x):
        out = np.zeros(2)
        lib().orc_conj_dot_z(_i64(x.size), _ptr(x), _ptr(y), _ptr(out))
        return complex(out[0], out[1])
    return float(lib().orc_conj_dot_d(_i64(x.size), _ptr(x), _ptr(y)))


def axpy(a, x, y) -> None:
    """y += a*x in place (vecalg.rs:571-575)."""
    assert y.flags.c_contiguous
    x = np.ascontiguousarray(x, dtype=y.dtype)
    if _sfx(y.dtype) == "s":
        lib().orc_axpy_s(_i64(y.size), C.c_float(a), _ptr(x), _ptr(y))
    elif _sfx(y.dtype) == "c":
        lib().orc_axpy_c(_i64(y.size), _ptr(_c2f(a)), _ptr(x), _ptr(y))
    elif _is_c(y):
        lib().orc_axpy_z(_i64(y.size), _ptr(_c2(a)), _ptr(x), _ptr(y))
    else:
        lib().orc_axpy_d(_i64(y.size), _dbl(a), _ptr(x), _ptr(y))


def axpby(a, x, b, y) -> None:
    """y = a*x + b*y in place (vecalg.rs:586-590)."""
    assert y.flags.c_contiguous
    x = np.ascontiguousarray(x, dtype=y.dtype)
    if _sfx(y.dtype) == "s":
        lib().orc_axpby_s(_i64(y.size), C.c_float(a), _ptr(x), C.c_float(b), _ptr(y))
    elif _sfx(y.dtype) == "c":
        lib().orc_axpby_c(_i64(y.size), _ptr(_c2f(a)), _ptr(x), _ptr(_c2f(b)), _ptr(y))
    elif _is_c(y):
        lib().orc_axpby_z(_i64(y.size), _ptr(_c2(a)), _ptr(x), _ptr(_c2(b)), _ptr(y))
    else:
        lib().orc_axpby_d(_i64(y.size), _dbl(a), _ptr(x), _dbl(b), _ptr(y))


def scale(a, x) -> None:
    assert x.flags.c_contiguous
    if _sfx(x.dtype) == "s":
        lib().orc_scale_s(_i64(x.size), C.c_float(a), _ptr(x))
    elif _sfx(x.dtype) == "c":
        lib().orc_scale_c(_i64(x.size), _ptr(_c2f(a)), _ptr(x))
    elif _is_c(x):
        lib().orc_scale_z(_i64(x.size), _ptr(_c2(a)), _ptr(x))
    else:
        lib().orc_scale_d(_i64(x.size), _dbl(a), _ptr(x))


def rscale(a: float, x) -> None:
    assert x.flags.c_contiguous
    if _is_single(x.dtype):
        getattr(lib(), f"orc_rscale_{_sfx(x.dtype)}")(_i64(x.size), C.c_float(a), _ptr(x))
    elif _is_c(x):
        lib().orc_rscale_z(_i64(x.size), _dbl(a), _ptr(x))
    else:
        lib().orc_scale_d(_i64(x.size), _dbl(a), _ptr(x))


def conj(x) -> np.ndarray:
    x = np.ascontiguousarray(x)
    if not _is_c(x):
        return x.copy()
    out = np.empty_like(x)
    getattr(lib(), f"orc_conj_{_sfx(x.dtype)}")(_i64(x.size), _ptr(x), _ptr(out))
    return out


# --------------------------------------------------------------------------- preconditioners
def diag_apply(diag, v) -> np.ndarray:
    """DiagPrecond::new(diag) then mul_vec (precond.rs:20-29, 48-52)."""
    v = np.ascontiguousarray(v)
    diag = np.ascontiguousarray(diag)
    out = np.empty_like(v)
    if _is_single(v.dtype):
        if _is_c(v) and not _is_c(diag):
            lib().orc_diag_apply_cs(_i64(v.size), _ptr(diag.astype(np.float32)), _ptr(v), _ptr(out))
        elif _is_c(v):
            lib().orc_diag_apply_c(_i64(v.size), _ptr(diag.astype(np.complex64)), _ptr(v), _ptr(out))
        else:
            lib().orc_diag_apply_s(_i64(v.size), _ptr(diag.astype(np.float32)), _ptr(v), _ptr(out))
        return out
    if _is_c(v) and not _is_c(diag):
        lib().orc_diag_apply_zd(_i64(v.size), _ptr(diag.astype(np.float64)), _ptr(v), _ptr(out))
    elif _is_c(v):
        lib().orc_diag_apply_z(_i64(v.size), _ptr(diag.astype(np.complex128)), _ptr(v), _ptr(out))
    else:
        lib().orc_diag_apply_d(_i64(v.size), _ptr(diag.astype(np.float64)), _ptr(v), _ptr(out))
    return out


def gs_apply(A: Csr, v, symmetric: bool, omega: float = 1.0) -> np.ndarray:
    """Gauss-Seidel operator; omega != 1: relaxed sweep (symmetric: SSOR(omega)), see sprs_oracle.h."""
    v = np.ascontiguousarray(v, dtype=A.dtype)
    out = np.empty_like(v)
    if omega != 1.0:
        st = getattr(lib(), f"orc_sor_apply_{_sfx(A.dtype)}")(
            _i64(A.n), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data), C.c_int(int(symmetric)), _dbl(omega), _ptr(v), _ptr(out)
        )
        if st != OK:
            raise ZeroDivisionError("zero diagonal")
        return out
    st = getattr(lib(), f"orc_gs_apply_{_sfx(A.dtype)}")(
        _i64(A.n), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data), C.c_int(int(symmetric)), _ptr(v), _ptr(out)
    )
    if st != OK:
        raise ZeroDivisionError("zero diagonal")
    return out


# --------------------------------------------------------------------------- solvers
@dataclass
class SolveOut:
    status: int
    iters: int
    resid: float
    x: np.ndarray
    hist: np.ndarray


def _pc_args(A: Csr, pc):
    """pc: None | ("diag", array) | ("gs_fwd",) | ("gs_sym",) | ("sor_fwd", omega) | ("ssor", omega)."""
    if pc is None:
        return PC_NONE, None
    kind = pc[0]
    if kind == "diag":
        d = np.ascontiguousarray(pc[1])
        if _is_c(A.data) and not _is_c(d):
            return PC_DIAG_REAL, d.astype(_real_dtype(A.dtype))
        return PC_DIAG, d.astype(A.dtype)
    if kind == "gs_fwd":
        return PC_GS_FWD, None
    if kind == "gs_sym":
        return PC_GS_SYM, None
    if kind in ("sor_fwd", "ssor"):
        return (PC_SOR_FWD if kind == "sor_fwd" else PC_SSOR), np.array([pc[1]], dtype=_real_dtype(A.dtype))
    raise ValueError(kind)


def _solve(name, nws, A: Csr, rhs, x0, max_iter, tol, pc, work, size, hist_cap, with_pc=True):
    rhs = np.ascontiguousarray(rhs, dtype=A.dtype)
    x = np.array(x0, dtype=A.dtype, copy=True) if x0 is not None else np.zeros(rhs.size, dtype=A.dtype)
    size = A.n if size is None else size
    if work is None:
        work = np.zeros(nws * size, dtype=A.dtype)
    hist = np.zeros(hist_cap, dtype=np.float64)
    iters = _i64(0)
    resid = _dbl(0.0)
    hlen = _i64(0)
    fn = getattr(lib(), f"orc_{name}_{_sfx(A.dtype)}")
    args = [_i64(size), _i64(rhs.size), _i64(x.size), _ptr(A.indptr), _ptr(A.indices), _ptr(A.data)]
    keep = None
    if with_pc:
        kind, keep = _pc_args(A, pc)
        args += [C.c_int(kind), _ptr(keep)]
    args += [
        _ptr(rhs), _ptr(x), _i64(max_iter), _dbl(tol), _ptr(work), C.byref(iters), C.byref(resid),
        _ptr(hist), _i64(hist_cap), C.byref(hlen),
    ]
    st = fn(*args)
    return SolveOut(int(st), int(iters.value), float(resid.value), x, hist[: min(hlen.value, hist_cap)].copy())


def bicgstab(A, rhs, x0=None, max_iter=1000, tol=1e-8, pc=None, work=None, size=None, hist_cap=4096):
    return _solve("bicgstab", 7, A, rhs, x0, max_iter, tol, pc, work, size, hist_cap)


def minres(A, rhs, x0=None, max_iter=1000, tol=1e-8, pc=None, work=None, size=None, hist_cap=4096):
    return _solve("minres", 8, A, rhs, x0, max_iter, tol, pc, work, size, hist_cap)


def csminres(A, rhs, x0=None, max_iter=1000, tol=1e-8, work=None, size=None, hist_cap=4096):
    return _solve("csminres", 7, A, rhs, x0, max_iter, tol, None, work, size, hist_cap, with_pc=False)


def gauss_seidel(A: Csr, rhs, x0=None, max_iter=300, eps=0.0, work=None, is_csr=True, hist_cap=4096, omega=1.0):
    rhs = np.ascontiguousarray(rhs, dtype=A.dtype)
    x = np.array(x0, dtype=A.dtype, copy=True) if x0 is not None else np.zeros(rhs.size, dtype=A.dtype)
    if work is None:
        work = np.zeros(2 * A.n, dtype=A.dtype)
    hist = np.zeros(hist_cap, dtype=np.float64)
    iters, resid, hlen = _i64(0), _dbl(0.0), _i64(0)
    if omega != 1.0:  # SOR: GaussSeidel::solve with the relaxed update (sprs_oracle.h)
        st = getattr(lib(), f"orc_sor_solve_{_sfx(A.dtype)}")(
            _i64(A.n), _i64(A.ncols), C.c_int(int(is_csr)), _i64(rhs.size), _i64(x.size), _ptr(A.indptr),
            _ptr(A.indices), _ptr(A.data), _ptr(rhs), _ptr(x), _i64(max_iter), _dbl(eps), _dbl(omega), _ptr(work),
            C.byref(iters), C.byref(resid), _ptr(hist), _i64(hist_cap), C.byref(hlen),
        )
        return SolveOut(int(st), int(iters.value), float(resid.value), x, hist[: min(hlen.value, hist_cap)].copy())
    st = getattr(lib(), f"orc_gauss_seidel_{_sfx(A.dtype)}")(
        _i64(A.n), _i64(A.ncols), C.c_int(int(is_csr)), _i64(rhs.size), _i64(x.size), _ptr(A.indptr),
        _ptr(A.indices), _ptr(A.data), _ptr(rhs), _ptr(x), _i64(max_iter), _dbl(eps), _ptr(work),
        C.byref(iters), C.byref(resid), _ptr(hist), _i64(hist_cap), C.byref(hlen),
    )
    return SolveOut(int(st), int(iters.value), float(resid.value), x, hist[: min(hlen.value, hist_cap)].copy())


# --------------------------------------------------------------------------- generators
def gen_dirichlet2d(rows: int, cols: int | None = None):
    """Reference generator src/main.rs:53-88 + rhs main.rs:90-103 (square grids)."""
    cols = rows if cols is None else cols
    n = rows * cols
    nnz = lib().orc_gen_dirichlet2d(_i64(rows), _i64(cols), None, None, None, None)
    indptr = np.empty(n + 1, np.int64)
    idx = np.empty(nnz, np.int32)
    a = np.empty(nnz, np.float64)
    rhs = np.zeros(n, np.float64)
    lib().orc_gen_dirichlet2d(_i64(rows), _i64(cols), _ptr(indptr), _ptr(idx), _ptr(a), _ptr(rhs))
    return Csr(n, indptr, idx, a), rhs


def gen_lap3d7(nx, ny=None, nz=None, shift=0.0, dtype=np.float64):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    if np.dtype(dtype).kind == "c":
        s = complex(shift)
        nnz = lib().orc_gen_lap3d7_z(_i64(nx), _i64(ny), _i64(nz), _dbl(s.real), _dbl(s.imag), None, None, None)
        indptr, idx, a = np.empty(n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.complex128)
        lib().orc_gen_lap3d7_z(_i64(nx), _i64(ny), _i64(nz), _dbl(s.real), _dbl(s.imag), _ptr(indptr), _ptr(idx), _ptr(a))
    else:
        nnz = lib().orc_gen_lap3d7_d(_i64(nx), _i64(ny), _i64(nz), _dbl(shift), None, None, None)
        indptr, idx, a = np.empty(n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
        lib().orc_gen_lap3d7_d(_i64(nx), _i64(ny), _i64(nz), _dbl(shift), _ptr(indptr), _ptr(idx), _ptr(a))
    return Csr(n, indptr, idx, a)


def gen_convdiff27(nx, ny=None, nz=None, b=(1.0, 0.5, 0.25), row_begin=0, row_end=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    row_end = n if row_end is None else row_end
    nloc = row_end - row_begin
    args = (_i64(nx), _i64(ny), _i64(nz), _dbl(b[0]), _dbl(b[1]), _dbl(b[2]), _i64(row_begin), _i64(row_end))
    nnz = lib().orc_gen_convdiff27_d(*args, None, None, None)
    indptr, idx, a = np.empty(nloc + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_gen_convdiff27_d(*args, _ptr(indptr), _ptr(idx), _ptr(a))
    return Csr(nloc, indptr, idx, a, ncols=n)
