"""vecalg (src/vecalg.rs:19-144) on host slices, computed on the GPU through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F
from .api import _check, _dtype_code, _ptr, default_context


def _prep(*arrs):
    """Common scalar type of the operands: f32 / Complex32 only when every operand is single precision."""
    arrs = [np.asarray(a) for a in arrs]
    single = all(a.dtype in (np.float32, np.complex64) for a in arrs)
    cplx = any(np.iscomplexobj(a) for a in arrs)
    dt = (np.complex64 if single else np.complex128) if cplx else (np.float32 if single else np.float64)
    return dt, [np.ascontiguousarray(a, dtype=dt) for a in arrs]


def _pair(a):
    a = complex(a)
    return (C.c_double * 2)(a.real, a.imag)


def _out_inplace(y, dt):
    if not (isinstance(y, np.ndarray) and y.dtype == dt and y.flags.c_contiguous):
        raise TypeError(f"in/out vector must be a contiguous {np.dtype(dt)} numpy array")
    return y


def _io_dtype(vec):
    """Scalar type of an in/out vector: its own if it is one of the four, else f64 / Complex64."""
    d = getattr(vec, "dtype", None)
    if d in (np.float32, np.complex64, np.float64, np.complex128):
        return np.dtype(d).type
    return np.complex128 if np.iscomplexobj(vec) else np.float64


def _scalar(out, dt):
    return complex(out[0], out[1]) if np.dtype(dt).kind == "c" else float(out[0])


def _h(ctx):
    return (ctx or default_context())._h


def dot(x, y, ctx=None):
    """x^T y, no conjugate (src/vecalg.rs:19-32)."""
    dt, (x, y) = _prep(x, y)
    assert x.size == y.size
    out = (C.c_double * 2)()
    _check(F.lib().spb_vec_dot(_h(ctx), _dtype_code(dt), x.size, _ptr(x), _ptr(y), out))
    return _scalar(out, dt)


def conj_dot(x, y, ctx=None):
    """x^H y, conjugate-linear in the first argument (src/vecalg.rs:34-59)."""
    dt, (x, y) = _prep(x, y)
    assert x.size == y.size
    out = (C.c_double * 2)()
    _check(F.lib().spb_vec_conj_dot(_h(ctx), _dtype_code(dt), x.size, _ptr(x), _ptr(y), out))
    return _scalar(out, dt)


def norm2(x, ctx=None) -> float:
    dt, (x,) = _prep(x)
    out = C.c_double(0.0)
    _check(F.lib().spb_vec_norm2(_h(ctx), _dtype_code(dt), x.size, _ptr(x), C.byref(out)))
    return float(out.value)


def scale(a, vec, ctx=None) -> None:
    dt = _io_dtype(vec)
    _out_inplace(vec, dt)
    _check(F.lib().spb_vec_scale(_h(ctx), _dtype_code(dt), vec.size, _pair(a), _ptr(vec)))


def rscale(a: float, vec, ctx=None) -> None:
    dt = _io_dtype(vec)
    _out_inplace(vec, dt)
    _check(F.lib().spb_vec_rscale(_h(ctx), _dtype_code(dt), vec.size, float(a), _ptr(vec)))


def conj(vec_in, vec_out, ctx=None) -> None:
    dt = _io_dtype(vec_out)
    _out_inplace(vec_out, dt)
    x = np.ascontiguousarray(vec_in, dtype=dt)
    assert x.size == vec_out.size
    _check(F.lib().spb_vec_conj(_h(ctx), _dtype_code(dt), x.size, _ptr(x), _ptr(vec_out)))


def axpy(a, vec1, vec2, ctx=None) -> None:
    """vec2 += a * vec1 (src/vecalg.rs:104-116)."""
    dt = _io_dtype(vec2)
    _out_inplace(vec2, dt)
    x = np.ascontiguousarray(vec1, dtype=dt)
    assert x.size == vec2.size
    _check(F.lib().spb_vec_axpy(_h(ctx), _dtype_code(dt), x.size, _pair(a), _ptr(x), _ptr(vec2)))


def axpby(a, vec1, b, vec2, ctx=None) -> None:
    """vec2 = a * vec1 + b * vec2 (src/vecalg.rs:118-144)."""
    dt = _io_dtype(vec2)
    _out_inplace(vec2, dt)
    x = np.ascontiguousarray(vec1, dtype=dt)
    assert x.size == vec2.size
    _check(F.lib().spb_vec_axpby(_h(ctx), _dtype_code(dt), x.size, _pair(a), _ptr(x), _pair(b), _ptr(vec2)))
