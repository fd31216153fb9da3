"""Host-side mirror of the reference's public interface for the iterative-solve path, on top of
the C ABI (include/sprsolve_b200.h).  Same names, argument meaning and error behaviour as the
Rust crate, so the parity tests read like the reference's own tests:

    reference (Rust)                               here
    ---------------------------------------------  -------------------------------------------
    trait MatVecMul<T>            src/mat.rs:12    class MatVecMul (mul_vec, mul_vec_dot, *_unchecked)
    MklMat::new(CsMatI<T,i32>)    src/mkl_mat.rs   GpuCsrMat.new(indptr, indices, data)
    DiagPrecond::new(&diag)       src/precond.rs   DiagPrecond.new(diag, dtype)
    BiCGStab::new(&A, n)          src/bicg_stab.rs BiCGStab(A, n).solve / .precond_solve
    MinRes / CSMinRes             src/minres.rs    MinRes / CSMinRes
    GaussSeidel::new(A.view())    src/gauss_seidel.rs GaussSeidel(A).solve
    SolverError::*                src/error.rs     SolverError subclasses
    vecalg::{dot,conj_dot,...}    src/vecalg.rs    sprsolve_b200.vecalg.*

Vectors are numpy arrays (float64 / complex128 / float32 / complex64) standing in for `&[T]` / `&mut [T]`; `x` is
updated in place like the reference's `&mut [T]`.  Everything computes on the GPU through the
C ABI; nothing here does arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi as F


# --------------------------------------------------------------------------- errors (src/error.rs)
class SolverError(Exception):
    pass


class IncompatibleMatrixFormat(SolverError):
    """SolverError::IncompatibleMatrixFormat(String)"""


class ZeorDiagonalElem(SolverError):  # sic, src/error.rs:11-12
    """SolverError::ZeorDiagonalElem(usize)"""

    def __init__(self, row):
        super().__init__(f"Matrix has zero diagonal element at {row}")
        self.row = row


class InsufficientIterNum(SolverError):
    """SolverError::InsufficientIterNum(usize)"""

    def __init__(self, n):
        super().__init__(f"Insufficient interation #: {n}")
        self.max_iter = n


class BreakDown(SolverError):
    """SolverError::BreakDown(usize)"""

    def __init__(self, its):
        super().__init__(f"Solver break down: its #{its}")
        self.its = its


class InvalidPreconditioner(SolverError):
    """SolverError::InvalidPreconditioner(String)"""


class DimensionMismatch(Exception):
    """panic!("Dimension mismatch") of mul_vec / mul_vec_dot (src/mat.rs:50-52)."""


class BackendError(RuntimeError):
    """CUDA / NCCL / argument failure inside libsprsolve_b200 (no reference analogue)."""


def _raise(status: int, iters: int = 0):
    msg = F.last_error()
    if status == F.INCOMPATIBLE_FORMAT:
        raise IncompatibleMatrixFormat(msg)
    if status == F.ZERO_DIAGONAL:
        raise ZeorDiagonalElem(iters)
    if status == F.INSUFFICIENT_ITER:
        raise InsufficientIterNum(iters)
    if status == F.BREAKDOWN:
        raise BreakDown(iters)
    if status == F.INVALID_PRECOND:
        raise InvalidPreconditioner(msg)
    if status == F.DIM_MISMATCH:
        raise DimensionMismatch("Dimension mismatch")
    if status == F.UNIMPLEMENTED:
        raise NotImplementedError(msg)
    raise BackendError(f"spb status {status}: {msg}")


def _check(status: int):
    if status != F.OK:
        _raise(status)


_CODES = {np.dtype(np.float64): F.F64, np.dtype(np.complex128): F.C128, np.dtype(np.float32): F.F32, np.dtype(np.complex64): F.C64}
SCALAR_TYPES = (np.float64, np.complex128, np.float32, np.complex64)  # cauchy::Scalar: f64, Complex64, f32, Complex32


def _dtype_code(dtype) -> int:
    k = np.dtype(dtype)
    if k in _CODES:
        return _CODES[k]
    raise TypeError(f"unsupported scalar type {k} (f32, f64, Complex<f32>, Complex<f64> are implemented)")


def _np_dtype(code: int):
    return {v: k for k, v in _CODES.items()}[code].type


def _as_scalar_array(data):
    """Contiguous array of one of the four scalar types; anything else is promoted to f64 / Complex64."""
    data = np.ascontiguousarray(data)
    if data.dtype not in _CODES:
        data = data.astype(np.complex128 if np.iscomplexobj(data) else np.float64)
    return data


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- context
class Context:
    """One GPU + one CUDA stream (+ optionally one NCCL rank)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(F.lib().spb_init(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            F.lib().spb_finalize(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        _check(F.lib().spb_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def synchronize(self):
        _check(F.lib().spb_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(F.lib().spb_launch_count(self._h))

    def profile(self, on: bool):
        _check(F.lib().spb_profile_enable(self._h, int(on)))

    def profile_read(self, family: int):
        n, ms = C.c_int64(0), C.c_double(0.0)
        _check(F.lib().spb_profile_read(self._h, family, C.byref(n), C.byref(ms)))
        return int(n.value), float(ms.value)

    def profile_reset(self):
        _check(F.lib().spb_profile_reset(self._h))

    # multi-GPU: one process per GPU, id created on rank 0 and broadcast by the caller
    @staticmethod
    def _prefer_torch_nccl():
        """The library resolves NCCL at run time (dlopen by soname).  When PyTorch is installed, load ITS copy
        first: two NCCL builds with the same soname cannot coexist in one process, and a torch imported
        after the system copy was mapped fails to resolve the newer symbols it was built against."""
        try:
            import torch  # noqa: F401
        except Exception:
            pass

    @staticmethod
    def comm_unique_id() -> bytes:
        Context._prefer_torch_nccl()
        buf = C.create_string_buffer(128)
        _check(F.lib().spb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, world: int, rank: int, unique_id: bytes):
        Context._prefer_torch_nccl()
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(F.lib().spb_comm_init(self._h, world, rank, buf))

    def comm_info(self):
        w, r = C.c_int(1), C.c_int(0)
        _check(F.lib().spb_comm_info(self._h, C.byref(w), C.byref(r)))
        return int(w.value), int(r.value)


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# --------------------------------------------------------------------------- trait MatVecMul<T>
class MatVecMul:
    """trait MatVecMul<T> (src/mat.rs:12-37)."""

    _h = None
    ctx: Context = None
    dtype = np.float64

    def _sizes(self):
        g, l, b = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(F.lib().spb_op_size(self._h, C.byref(g), C.byref(l), C.byref(b)))
        return int(g.value), int(l.value), int(b.value)

    def size(self) -> int:
        """MklMat::size (src/mkl_mat.rs:26-28): the (global) dimension."""
        return self._sizes()[0]

    @property
    def n_local(self) -> int:
        return self._sizes()[1]

    @property
    def row_begin(self) -> int:
        return self._sizes()[2]

    def _vec(self, v, writable=False):
        a = np.asarray(v)
        if a.dtype != self.dtype or not a.flags.c_contiguous:
            if writable:
                raise TypeError(f"output vector must be a contiguous {np.dtype(self.dtype)} array")
            a = np.ascontiguousarray(a, dtype=self.dtype)
        return a

    def mul_vec(self, v_in, v_out) -> None:
        """v_out = A v_in; raises DimensionMismatch like the reference panics (src/mat.rs:49-56)."""
        a, b = self._vec(v_in), self._vec(v_out, True)
        _check(F.lib().spb_op_mul_vec(self._h, _ptr(a), a.size, _ptr(b), b.size))

    def mul_vec_dot(self, v_in, v_out):
        """v_out = A v_in, returns conj(v_in) . v_out (src/mat.rs:59-64)."""
        a, b = self._vec(v_in), self._vec(v_out, True)
        out = (C.c_double * 2)()
        _check(F.lib().spb_op_mul_vec_dot(self._h, _ptr(a), a.size, _ptr(b), b.size, out))
        return complex(out[0], out[1]) if np.dtype(self.dtype).kind == "c" else float(out[0])

    # The reference's `unsafe fn *_unchecked` skip the dimension test; a wrong size there is
    # undefined behaviour, so the checked entry point is the only sensible Python mapping.
    mul_vec_unchecked = mul_vec
    mul_vec_dot_unchecked = mul_vec_dot

    def mul_vec_dev(self, d_in: int, d_out: int) -> None:
        _check(F.lib().spb_op_mul_vec_dev(self._h, C.c_void_p(d_in), C.c_void_p(d_out)))

    def mul_vec_dot_dev(self, d_in: int, d_out: int):
        out = (C.c_double * 2)()
        _check(F.lib().spb_op_mul_vec_dot_dev(self._h, C.c_void_p(d_in), C.c_void_p(d_out), out))
        return complex(out[0], out[1]) if np.dtype(self.dtype).kind == "c" else float(out[0])

    def destroy(self):
        if self._h:
            F.lib().spb_op_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class GpuCsrMat(MatVecMul):
    """GPU-resident CSR matrix: the drop-in for MklMat<T> (src/mkl_mat.rs:15-149) and for the
    CsMatI<T,I> operator (src/mat.rs:47-183)."""

    def __init__(self, handle, ctx, dtype):
        self._h, self.ctx, self.dtype = handle, ctx, dtype

    @classmethod
    def new(cls, indptr, indices, data, shape=None, ctx: Context | None = None, row_range=None):
        """MklMat::new(m): takes the raw CSR storage (into_raw_storage, src/mkl_mat.rs:41).
        Must be square CSR (src/mkl_mat.rs:33-37) else IncompatibleMatrixFormat.
        row_range=(begin, end): this rank's row block of a partitioned matrix (indptr local,
        indices global)."""
        ctx = ctx or default_context()
        data = _as_scalar_array(data)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr)
        if indptr.dtype == np.int32:
            bits = 32
        else:
            indptr = indptr.astype(np.int64)
            bits = 64
        nloc = indptr.size - 1
        if shape is None:
            shape = (nloc, nloc)
        rb, re = row_range if row_range is not None else (0, shape[0])
        h = C.c_void_p()
        st = F.lib().spb_csr_create(
            ctx._h, _dtype_code(data.dtype), shape[0], shape[1], rb, re, _ptr(indptr), bits, _ptr(indices), _ptr(data), C.byref(h)
        )
        _check(st)
        return cls(h, ctx, data.dtype.type)

    @staticmethod
    def _values(data):
        return _as_scalar_array(data)

    @classmethod
    def from_csc(cls, indptr, row_indices, data, shape=None, ctx: Context | None = None):
        """A CSC matrix as the operator (the CSC branch of CsMatViewI::mul_vec, src/mat.rs:130-142):
        transposed on the device with a stable sort, so every output element keeps the accumulation
        order of the reference's column-by-column loop."""
        ctx = ctx or default_context()
        data = cls._values(data)
        row_indices = np.ascontiguousarray(row_indices, dtype=np.int32)
        indptr = np.ascontiguousarray(indptr)
        bits = 32 if indptr.dtype == np.int32 else 64
        if bits == 64:
            indptr = indptr.astype(np.int64)
        n = indptr.size - 1
        shape = shape or (n, n)
        h = C.c_void_p()
        _check(F.lib().spb_csc_create(ctx._h, _dtype_code(data.dtype), shape[0], shape[1], _ptr(indptr), bits, _ptr(row_indices), _ptr(data), C.byref(h)))
        return cls(h, ctx, data.dtype.type)

    @classmethod
    def from_triplets(cls, n: int, rows, cols, data, ctx: Context | None = None):
        """sprs::TriMat::to_csr (tests/test_minres.rs:65-119): sorted by (row, column), duplicates
        summed in input order; assembled on the device."""
        ctx = ctx or default_context()
        data = cls._values(data)
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        if not (rows.size == cols.size == data.size):
            raise ValueError("rows, cols and data must have the same length")
        h = C.c_void_p()
        _check(F.lib().spb_csr_create_from_triplets(ctx._h, _dtype_code(data.dtype), n, n, rows.size, _ptr(rows), _ptr(cols), _ptr(data), C.byref(h)))
        return cls(h, ctx, data.dtype.type)

    @classmethod
    def read_matrix_market(cls, path: str, dtype=np.float64, ctx: Context | None = None):
        """Matrix Market coordinate file -> device CSR (spb_csr_read_matrix_market)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        _check(F.lib().spb_csr_read_matrix_market(ctx._h, _dtype_code(dtype), os.fsencode(path), C.byref(h)))
        return cls(h, ctx, np.dtype(dtype).type)

    @classmethod
    def from_stencil(cls, kind: int, nx: int, ny: int | None = None, nz: int | None = None, params=(), dtype=np.float64, ctx=None):
        """On-device synthetic generator; the row block of this rank when
        the context has a communicator."""
        ctx = ctx or default_context()
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        p = (C.c_double * max(len(params), 1))(*[float(v) for v in params])
        h = C.c_void_p()
        _check(F.lib().spb_csr_create_stencil(ctx._h, kind, _dtype_code(dtype), nx, ny, nz, p, len(params), C.byref(h)))
        return cls(h, ctx, np.dtype(dtype).type)

    def mv_hint(self, ncalls: int = 2000):
        """MklMat::mv_hint (src/mkl_mat.rs:124-148)."""
        _check(F.lib().spb_csr_mv_hint(self._h, ncalls))

    def mv_and_dotmv_hint(self, ncalls: int = 2000):
        """MklMat::mv_and_dotmv_hint (src/mkl_mat.rs:81-118)."""
        _check(F.lib().spb_csr_mv_and_dotmv_hint(self._h, ncalls))

    @property
    def nnz(self) -> int:
        v = C.c_int64(0)
        _check(F.lib().spb_csr_nnz(self._h, C.byref(v)))
        return int(v.value)

    def plan_info(self) -> dict:
        """What the analysis decided (spb_csr_plan_info)."""
        v = (C.c_int64 * 8)()
        _check(F.lib().spb_csr_plan_info(self._h, v))
        keys = ["dictionary", "patterns", "dict_width", "consumer_threads", "stages", "tile_nnz", "ctas_per_sm", "stream_bytes"]
        d = {k: int(v[i]) for i, k in enumerate(keys)}
        d["x_window"] = (d["dictionary"] >> 1) & 1  # x segments staged in shared memory by bulk copies
        d["dictionary"] &= 1
        return d

    def download(self):
        """(indptr int64, indices int32 global, data) of the local rows."""
        n, nnz = self.n_local, self.nnz
        indptr = np.empty(n + 1, np.int64)
        idx = np.empty(nnz, np.int32)
        data = np.empty(nnz, self.dtype)
        _check(F.lib().spb_csr_download(self._h, _ptr(indptr), _ptr(idx), _ptr(data)))
        return indptr, idx, data

    def diagonal(self) -> np.ndarray:
        d = np.empty(self.n_local, self.dtype)
        _check(F.lib().spb_csr_diagonal(self._h, _ptr(d)))
        return d


class DiagPrecond(MatVecMul):
    """DiagPrecond<T, V> (src/precond.rs:6-63): Jacobi; V may be real while T is complex."""

    def __init__(self, handle, ctx, dtype):
        self._h, self.ctx, self.dtype = handle, ctx, dtype

    @classmethod
    def new(cls, diag, dtype=None, ctx: Context | None = None):
        """DiagPrecond::new(diag): stores 1/diag (src/precond.rs:20-29)."""
        ctx = ctx or default_context()
        diag = _as_scalar_array(diag)
        ddt = diag.dtype.type
        dtype = np.dtype(dtype or ddt).type
        # V = T or V = T::Real: bring the diagonal to the precision of the system
        single = np.dtype(dtype).itemsize in (4, 8) and np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
        want = (np.complex64 if single else np.complex128) if np.iscomplexobj(diag) else (np.float32 if single else np.float64)
        if ddt != want:
            diag = diag.astype(want)
            ddt = want
        h = C.c_void_p()
        _check(F.lib().spb_diag_precond_create(ctx._h, _dtype_code(dtype), _dtype_code(ddt), _ptr(diag), diag.size, C.byref(h)))
        return cls(h, ctx, dtype)

    @classmethod
    def from_matrix(cls, A: GpuCsrMat):
        h = C.c_void_p()
        _check(F.lib().spb_diag_precond_from_csr(A._h, C.byref(h)))
        return cls(h, A.ctx, A.dtype)


class GaussSeidelPrecond(MatVecMul):
    """Level-scheduled Gauss-Seidel sweep as a MatVecMul operator (sweep body
    src/gauss_seidel.rs:111-125; the reference has no such wrapper -- see DESIGN.md)."""

    def __init__(self, A: GpuCsrMat, symmetric: bool = False, omega: float = 1.0):
        """omega != 1: the relaxed sweep (SOR; symmetric: SSOR(omega)), spb_gs_precond_create_relaxed."""
        h = C.c_void_p()
        st = F.lib().spb_gs_precond_create_relaxed(A._h, F.GS_SYMMETRIC if symmetric else F.GS_FORWARD, float(omega), C.byref(h))
        if st == F.ZERO_DIAGONAL:
            msg = F.last_error()
            raise ZeorDiagonalElem(int(msg.rsplit(" ", 1)[-1]))
        _check(st)
        self._h, self.ctx, self.dtype, self._A, self.omega = h, A.ctx, A.dtype, A, float(omega)

    def levels(self):
        a, b = C.c_int64(0), C.c_int64(0)
        _check(F.lib().spb_gs_levels(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def schedule_info(self, stats_blocks: int = 0):
        """Diagnostics of the block-wavefront sweep schedule (spb_gs_schedule_info)."""
        info = (C.c_int64 * 16)()
        st = (C.c_int64 * max(4 * stats_blocks, 1))()
        _check(F.lib().spb_gs_schedule_info(self._h, info, st if stats_blocks else None, 4 * stats_blocks))
        keys = ["fwd_ok", "block_rows", "blocks", "fwd_chunks", "fwd_local_levels", "stages", "smem_bytes", "bwd_ok",
                "bwd_chunks", "bwd_local_levels", "poll_timeout", "rhs_slots", "other_slots", "packed_bytes"]
        d = {k: int(info[i]) for i, k in enumerate(keys)}
        if stats_blocks:
            d["stats"] = np.array(st[: 4 * stats_blocks], dtype=np.int64).reshape(-1, 4)
        return d


# --------------------------------------------------------------------------- solvers
class _Solver:
    _create = None
    _needs_size = True

    def __init__(self, A: MatVecMul, size: int | None = None):
        self.A = A
        self.ctx = A.ctx
        self.dtype = A.dtype
        self.history = np.zeros(0)
        self.hist_cap = 0
        h = C.c_void_p()
        fn = getattr(F.lib(), self._create)
        st = fn(A._h, C.byref(h)) if not self._needs_size else fn(A._h, int(size if size is not None else A.n_local), C.byref(h))
        _check(st)
        self._h = h

    def record_history(self, capacity: int):
        """Keep the per-iteration residual of the next solves in `self.history`."""
        self.hist_cap = int(capacity)
        return self

    def set_poll_interval(self, iters: int):
        _check(F.lib().spb_solver_set_poll_interval(self._h, int(iters)))
        return self

    def _hist(self):
        return np.zeros(max(self.hist_cap, 1), np.float64) if self.hist_cap else None

    def _finish(self, st, iters, resid, hist, hlen):
        self.history = hist[: min(hlen.value, self.hist_cap)].copy() if hist is not None else np.zeros(0)
        if st != F.OK:
            _raise(st, int(iters.value))
        return int(iters.value), float(resid.value)

    def _solve(self, precond, rhs, x, max_iter, tol):
        rhs = np.ascontiguousarray(rhs, dtype=self.dtype)
        if not (isinstance(x, np.ndarray) and x.dtype == self.dtype and x.flags.c_contiguous):
            raise TypeError(f"x must be a contiguous {np.dtype(self.dtype)} numpy array (it is updated in place)")
        iters, resid, hlen = C.c_int64(0), C.c_double(0.0), C.c_int64(0)
        hist = self._hist()
        st = F.lib().spb_solver_solve(
            self._h, precond._h if precond is not None else None, _ptr(rhs), rhs.size, _ptr(x), x.size,
            int(max_iter), float(tol), C.byref(iters), C.byref(resid), _ptr(hist), self.hist_cap, C.byref(hlen),
        )
        return self._finish(st, iters, resid, hist, hlen)

    def _solve_dev(self, precond, d_rhs: int, d_x: int, max_iter, tol):
        iters, resid, hlen = C.c_int64(0), C.c_double(0.0), C.c_int64(0)
        hist = self._hist()
        st = F.lib().spb_solver_solve_dev(
            self._h, precond._h if precond is not None else None, C.c_void_p(d_rhs), C.c_void_p(d_x),
            int(max_iter), float(tol), C.byref(iters), C.byref(resid), _ptr(hist), self.hist_cap, C.byref(hlen),
        )
        return self._finish(st, iters, resid, hist, hlen)

    def solve(self, rhs, x, max_iter: int, tol: float):
        """solve(&mut self, rhs, x, max_iter, tol) -> Ok((iters, resid)) or raises SolverError."""
        return self._solve(None, rhs, x, max_iter, tol)

    def solve_dev(self, d_rhs: int, d_x: int, max_iter: int, tol: float, precond=None):
        """Same with rhs / x resident in device memory (raw device pointers)."""
        return self._solve_dev(precond, d_rhs, d_x, max_iter, tol)

    def destroy(self):
        if getattr(self, "_h", None):
            F.lib().spb_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class BiCGStab(_Solver):
    """BiCGStab<'data, T, M> (src/bicg_stab.rs:17-367)."""

    _create = "spb_bicgstab_create"

    def precond_solve(self, precond: MatVecMul, rhs, x, max_iter: int, tol: float):
        return self._solve(precond, rhs, x, max_iter, tol)


class MinRes(_Solver):
    """MinRes<'data, T, M> (src/minres.rs:13-342); the preconditioner must be SPD (:176)."""

    _create = "spb_minres_create"

    def precond_solve(self, precond: MatVecMul, rhs, x, max_iter: int, tol: float):
        return self._solve(precond, rhs, x, max_iter, tol)


class CSMinRes(_Solver):
    """CSMinRes<'data, T, M> (src/cs_minres.rs:11-159): complex-symmetric MINRES."""

    _create = "spb_csminres_create"


class GaussSeidel(_Solver):
    """GaussSeidel<'data, T> (src/gauss_seidel.rs:8-141); returns the ABSOLUTE residual."""

    _create = "spb_gauss_seidel_create"
    _needs_size = False

    def __init__(self, A: MatVecMul, omega: float = 1.0):
        """omega != 1: successive over-relaxation (spb_gauss_seidel_create_relaxed); 1: the reference's solver."""
        if omega == 1.0:
            super().__init__(A)
            return
        self.A, self.ctx, self.dtype = A, A.ctx, A.dtype
        self.history, self.hist_cap = np.zeros(0), 0
        h = C.c_void_p()
        _check(F.lib().spb_gauss_seidel_create_relaxed(A._h, float(omega), C.byref(h)))
        self._h = h

    def solve(self, rhs, x, max_iter: int, eps: float):
        return self._solve(None, rhs, x, max_iter, eps)
