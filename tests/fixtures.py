"""Matrix fixtures restating the generators of the reference's own tests (pure numpy, small sizes).

Each function cites the reference test file it mirrors (paths relative to the reference crate).
sprs::TriMat::to_csr sorts column indices within a row and sums duplicates; none of these
generators emits duplicates, so sorting the triplets row-major reproduces its output.
"""
from __future__ import annotations

import numpy as np

from oracle.oracle import Csr


def _from_triplets(n, trip, dtype):
    trip.sort(key=lambda t: (t[0], t[1]))
    indptr = np.zeros(n + 1, np.int64)
    for r, _, _ in trip:
        indptr[r + 1] += 1
    indptr = np.cumsum(indptr)
    idx = np.array([c for _, c, _ in trip], np.int32)
    data = np.array([v for _, _, v in trip], dtype)
    return Csr(n, indptr, idx, data)


# ---- src/mat.rs:232-255 / src/mkl_mat.rs:342-365 : 5x5 CSR with an empty row ----------------
KAT_INDPTR = [0, 3, 3, 5, 6, 7]
KAT_INDICES = [1, 2, 3, 2, 3, 4, 4]
KAT_DATA = [0.75672424, 0.1649078, 0.30140296, 0.10358244, 0.6283315, 0.39244208, 0.57202407]
KAT_X = [0.1, 0.2, -0.1, 0.3, 0.9]
KAT_Y = [0.22527496, 0.0, 0.17814121, 0.35319787, 0.51482166]

# ---- src/mat.rs:208-229 : 5x5 CSC -----------------------------------------------------------
KAT_CSC_INDPTR = [0, 2, 4, 5, 6, 7]
KAT_CSC_INDICES = [2, 3, 3, 4, 2, 1, 3]
KAT_CSC_DATA = [0.35310881, 0.42380633, 0.28035896, 0.58082095, 0.53350123, 0.88132896, 0.72527863]
KAT_CSC_Y = [0.0, 0.26439869, -0.01803924, 0.75120319, 0.11616419]

# ---- src/mkl_mat.rs:408-430 : 13-nnz integer-valued matrix, exact to 1e-16 -------------------
KAT2_INDPTR = [0, 3, 5, 8, 11, 13]
KAT2_INDICES = [0, 1, 3, 0, 1, 2, 3, 4, 0, 2, 3, 1, 4]
KAT2_DATA = [1.0, -1.0, -3.0, -2.0, 5.0, 4.0, 6.0, 4.0, -4.0, 2.0, 7.0, 8.0, -5.0]
KAT2_X = [1.0, 5.0, 1.0, 4.0, 1.0]
KAT2_Y = [-16.0, 23.0, 32.0, 26.0, 35.0]


def kat_csr(dtype=np.float64) -> Csr:
    data = np.array(KAT_DATA, np.float64)
    if np.dtype(dtype).kind == "c":  # src/mkl_mat.rs:372-380: v + v i
        data = data + 1j * data
    return Csr(5, np.array(KAT_INDPTR), np.array(KAT_INDICES), data)


def kat2_csr() -> Csr:
    return Csr(5, np.array(KAT2_INDPTR), np.array(KAT2_INDICES), np.array(KAT2_DATA))


def sym_laplacian_2d(rows, cols):
    """tests/test_minres.rs:76-120 -- symmetric 5-point grid (-4 diag, +1 off), boundary values
    bv(row,col)=row+col folded into the rhs."""
    n = rows * cols
    rhs = np.zeros(n)
    trip = []
    bv = lambda r, c: float(r + c)
    for i in range(rows):
        for j in range(cols):
            vid = i * cols + j
            trip.append((vid, vid, -4.0))
            if i > 0:
                trip.append((vid, (i - 1) * cols + j, 1.0))
            else:
                rhs[vid] -= bv(i - 1, j)
            if j > 0:
                trip.append((vid, i * cols + j - 1, 1.0))
            else:
                rhs[vid] -= bv(i, j - 1)
            if i < rows - 1:
                trip.append((vid, (i + 1) * cols + j, 1.0))
            else:
                rhs[vid] -= bv(i + 1, j)
            if j < cols - 1:
                trip.append((vid, i * cols + j + 1, 1.0))
            else:
                rhs[vid] -= bv(i, j + 1)
    return _from_triplets(n, trip, np.float64), rhs


def diag_simple(rows, cols):
    """tests/test_minres.rs:62-74 -- diag 2(i+1), rhs i+1."""
    n = rows * cols
    trip = [(i, i, float((i + 1) * 2)) for i in range(n)]
    return _from_triplets(n, trip, np.float64), np.arange(1, n + 1, dtype=np.float64)


def _complex_grid(rows, cols, diag_fn, off_fn):
    n = rows * cols
    rhs = np.zeros(n, np.complex128)
    trip = []
    diag = []
    val = lambda r, c: complex(r, c)  # known solution x*[i,j] = i + j*1i
    for i in range(rows):
        for j in range(cols):
            vid = i * cols + j
            c = diag_fn(i, j)
            diag.append(c)
            trip.append((vid, vid, c))
            rv = c * val(i, j)
            for (ii, jj, ok) in ((i - 1, j, i > 0), (i, j - 1, j > 0), (i + 1, j, i < rows - 1), (i, j + 1, j < cols - 1)):
                if ok:
                    tid = ii * cols + jj
                    cc = off_fn(vid, tid)
                    trip.append((vid, tid, cc))
                    rv += cc * val(ii, jj)
            rhs[vid] = rv
    xstar = np.array([val(i, j) for i in range(rows) for j in range(cols)], np.complex128)
    return _from_triplets(n, trip, np.complex128), rhs, np.array(diag), xstar


def hermitian_grid(rows, cols):
    """tests/test_complex_solve.rs:95-214 -- Hermitian grid: diag -3-i (real), off-diagonal
    1+2.5i below / 1-2.5i above the diagonal.  Returns (A, rhs, real_diag = -re(a_ii), x*)."""
    A, rhs, diag, xs = _complex_grid(
        rows, cols, lambda i, j: complex(-3.0 - i, 0.0), lambda r, c: complex(1.0, 2.5) if r > c else complex(1.0, -2.5)
    )
    return A, rhs, (-diag.real).astype(np.float64), xs


def complex_symmetric_grid(rows, cols):
    """tests/test_complex_solve2.rs:35-96 -- complex-symmetric grid: diag (-2-i)+(-2-j)i, all
    off-diagonals 1-2.5i.  Returns (A, rhs, complex diag, x*)."""
    return _complex_grid(rows, cols, lambda i, j: complex(-2.0 - i, -2.0 - j), lambda r, c: complex(1.0, -2.5))
