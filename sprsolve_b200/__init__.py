"""sprsolve_b200 -- B200 (sm_100a) drop-in for the iterative-solve hot path of cxzheng/sprsolve.

The product is the C-ABI library sprsolve_b200/lib/libsprsolve_b200.so (include/sprsolve_b200.h);
this package is the host-side mirror of the reference's Rust interface on top of it.  There is
no CPU fallback: importing works anywhere, but every operation needs the built library and a GPU.
"""
from . import _ffi
from .api import (
    BackendError,
    BiCGStab,
    BreakDown,
    Context,
    CSMinRes,
    DiagPrecond,
    DimensionMismatch,
    GaussSeidel,
    GaussSeidelPrecond,
    GpuCsrMat,
    IncompatibleMatrixFormat,
    InsufficientIterNum,
    InvalidPreconditioner,
    MatVecMul,
    MinRes,
    SolverError,
    ZeorDiagonalElem,
    default_context,
)
from . import vecalg

STENCIL_DIRICHLET2D = _ffi.STENCIL_DIRICHLET2D
STENCIL_LAP3D7 = _ffi.STENCIL_LAP3D7
STENCIL_CONVDIFF27 = _ffi.STENCIL_CONVDIFF27
