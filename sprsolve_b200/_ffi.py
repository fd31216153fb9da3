"""ctypes binding of include/sprsolve_b200.h (the C-ABI drop-in boundary).

Loading fails loudly when the shared library is missing: there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPB_LIB") or os.path.join(_HERE, "lib", "libsprsolve_b200.so")  # SPB_LIB: A/B builds (tools/)

# spb_status (include/sprsolve_b200.h)
OK = 0
INCOMPATIBLE_FORMAT = 1
ZERO_DIAGONAL = 2
INSUFFICIENT_ITER = 3
BREAKDOWN = 4
INVALID_PRECOND = 5
DIM_MISMATCH = 6
UNIMPLEMENTED = 7
CUDA_ERROR = 100
NCCL_ERROR = 101
INVALID_ARG = 102
NO_DEVICE = 103

F64, C128, F32, C64 = 0, 1, 2, 3
STENCIL_DIRICHLET2D, STENCIL_LAP3D7, STENCIL_CONVDIFF27 = 0, 1, 2
GS_FORWARD, GS_SYMMETRIC = 0, 1

i64 = C.c_int64
dbl = C.c_double
vp = C.c_void_p
pp = C.POINTER(C.c_void_p)
pi64 = C.POINTER(C.c_int64)
pdbl = C.POINTER(C.c_double)

# name -> (restype, argtypes): exactly the symbols include/sprsolve_b200.h declares
SIGNATURES = {
    "spb_version": (C.c_char_p, []),
    "spb_last_error": (C.c_char_p, []),
    "spb_init": (C.c_int, [C.c_int, pp]),
    "spb_finalize": (C.c_int, [vp]),
    "spb_set_stream": (C.c_int, [vp, vp]),
    "spb_synchronize": (C.c_int, [vp]),
    "spb_launch_count": (i64, [vp]),
    "spb_profile_enable": (C.c_int, [vp, C.c_int]),
    "spb_profile_read": (C.c_int, [vp, C.c_int, pi64, pdbl]),
    "spb_profile_reset": (C.c_int, [vp]),
    "spb_comm_unique_id": (C.c_int, [vp]),
    "spb_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "spb_comm_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "spb_stencil_partition": (C.c_int, [C.c_int, i64, i64, i64, C.c_int, C.c_int, pi64, pi64]),
    "spb_csr_create": (C.c_int, [vp, C.c_int, i64, i64, i64, i64, vp, C.c_int, vp, vp, pp]),
    "spb_csr_create_stencil": (C.c_int, [vp, C.c_int, C.c_int, i64, i64, i64, pdbl, C.c_int, pp]),
    "spb_csr_plan_info": (C.c_int, [vp, pi64]),
    "spb_csc_create": (C.c_int, [vp, C.c_int, i64, i64, vp, C.c_int, vp, vp, pp]),
    "spb_csr_create_from_triplets": (C.c_int, [vp, C.c_int, i64, i64, i64, vp, vp, vp, pp]),
    "spb_csr_read_matrix_market": (C.c_int, [vp, C.c_int, C.c_char_p, pp]),
    "spb_csr_mv_hint": (C.c_int, [vp, C.c_int]),
    "spb_csr_mv_and_dotmv_hint": (C.c_int, [vp, C.c_int]),
    "spb_op_size": (C.c_int, [vp, pi64, pi64, pi64]),
    "spb_csr_nnz": (C.c_int, [vp, pi64]),
    "spb_csr_download": (C.c_int, [vp, vp, vp, vp]),
    "spb_csr_diagonal": (C.c_int, [vp, vp]),
    "spb_op_destroy": (C.c_int, [vp]),
    "spb_op_mul_vec": (C.c_int, [vp, vp, i64, vp, i64]),
    "spb_op_mul_vec_dot": (C.c_int, [vp, vp, i64, vp, i64, pdbl]),
    "spb_op_mul_vec_dev": (C.c_int, [vp, vp, vp]),
    "spb_op_mul_vec_dot_dev": (C.c_int, [vp, vp, vp, pdbl]),
    "spb_diag_precond_create": (C.c_int, [vp, C.c_int, C.c_int, vp, i64, pp]),
    "spb_diag_precond_from_csr": (C.c_int, [vp, pp]),
    "spb_gs_precond_create": (C.c_int, [vp, C.c_int, pp]),
    "spb_gs_precond_create_relaxed": (C.c_int, [vp, C.c_int, dbl, pp]),
    "spb_gs_levels": (C.c_int, [vp, pi64, pi64]),
    "spb_gs_schedule_info": (C.c_int, [vp, pi64, pi64, i64]),
    "spb_vec_dot": (C.c_int, [vp, C.c_int, i64, vp, vp, pdbl]),
    "spb_vec_conj_dot": (C.c_int, [vp, C.c_int, i64, vp, vp, pdbl]),
    "spb_vec_norm2": (C.c_int, [vp, C.c_int, i64, vp, pdbl]),
    "spb_vec_scale": (C.c_int, [vp, C.c_int, i64, pdbl, vp]),
    "spb_vec_rscale": (C.c_int, [vp, C.c_int, i64, dbl, vp]),
    "spb_vec_conj": (C.c_int, [vp, C.c_int, i64, vp, vp]),
    "spb_vec_axpy": (C.c_int, [vp, C.c_int, i64, pdbl, vp, vp]),
    "spb_vec_axpby": (C.c_int, [vp, C.c_int, i64, pdbl, vp, pdbl, vp]),
    "spb_bicgstab_create": (C.c_int, [vp, i64, pp]),
    "spb_minres_create": (C.c_int, [vp, i64, pp]),
    "spb_csminres_create": (C.c_int, [vp, i64, pp]),
    "spb_gauss_seidel_create": (C.c_int, [vp, pp]),
    "spb_gauss_seidel_create_relaxed": (C.c_int, [vp, dbl, pp]),
    "spb_solver_solve": (C.c_int, [vp, vp, vp, i64, vp, i64, i64, dbl, pi64, pdbl, vp, i64, pi64]),
    "spb_solver_solve_dev": (C.c_int, [vp, vp, vp, vp, i64, dbl, pi64, pdbl, vp, i64, pi64]),
    "spb_solver_set_poll_interval": (C.c_int, [vp, C.c_int]),
    "spb_solver_destroy": (C.c_int, [vp]),
}

_lib = None


def lib() -> C.CDLL:
    """The C-ABI library.  Raises if it has not been built (python -m sprsolve_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m sprsolve_b200.build` "
                "(sprsolve_b200 has no CPU fallback)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().spb_last_error().decode("utf-8", "replace")
