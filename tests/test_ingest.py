"""Matrix ingestion on the callers' side of the operator boundary:
CSC operator (src/mat.rs:130-142, KAT :208-229), triplet assembly (sprs::TriMat::to_csr as used
by tests/test_minres.rs:65-119) and Matrix Market files.  GPU tests, through the C ABI."""
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sps

import fixtures as fx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import sprsolve_b200 as sp

    sp.default_context()
    return sp


def test_csc_kat(sp, orc):
    """src/mat.rs:208-229: the reference's CSC known-answer test, eps 1e-8; and bit-identical to the
    restated column-by-column loop."""
    G = sp.GpuCsrMat.from_csc(np.array(fx.KAT_CSC_INDPTR, np.int32), fx.KAT_CSC_INDICES, fx.KAT_CSC_DATA)
    y = np.zeros(5)
    G.mul_vec(np.array(fx.KAT_X), y)
    assert np.all(np.abs(y - np.array(fx.KAT_CSC_Y)) < 1e-8)
    assert np.array_equal(y, orc.spmv_csc(5, 5, fx.KAT_CSC_INDPTR, fx.KAT_CSC_INDICES, fx.KAT_CSC_DATA, fx.KAT_X))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_csc_operator_bit_exact(sp, orc, dtype):
    """A random CSC matrix with unsorted rows inside the columns: mul_vec == the reference's CSC loop
    (src/mat.rs:134-141), bit for bit."""
    rng = np.random.default_rng(7)
    n = 300
    M = sps.random(n, n, density=0.03, random_state=11, format="csc", dtype=np.float64)
    data = M.data.astype(dtype)
    if dtype is np.complex128:
        data = data + 1j * rng.uniform(-1, 1, data.size)
    indptr, idx = M.indptr.astype(np.int64), M.indices.astype(np.int32)
    for c in range(n):  # shuffle the rows inside every column: CSC storage need not be sorted
        s, e = indptr[c], indptr[c + 1]
        p = rng.permutation(e - s)
        idx[s:e], data[s:e] = idx[s:e][p], data[s:e][p]
    x = rng.uniform(-1, 1, n).astype(dtype)
    G = sp.GpuCsrMat.from_csc(indptr, idx, data)
    y = np.zeros(n, dtype)
    G.mul_vec(x, y)
    assert np.array_equal(y, orc.spmv_csc(n, n, indptr, idx, data, x))


def test_triplets_match_sprs_to_csr(sp, orc):
    """tests/test_minres.rs:80-119 builds its matrix with TriMat::add_triplet + to_csr; the device
    assembly gives the same CSR as the row-major restatement in tests/fixtures.py, and sums
    duplicates."""
    A, rhs = fx.minres_grid(6, 5) if hasattr(fx, "minres_grid") else (orc.gen_lap3d7(6, 5, 4, shift=0.3), None)
    S = A.to_scipy().tocoo()
    rng = np.random.default_rng(3)
    p = rng.permutation(S.nnz)  # triplets in arbitrary order
    G = sp.GpuCsrMat.from_triplets(A.n, S.row[p], S.col[p], S.data[p])
    ip, idx, dat = G.download()
    assert np.array_equal(ip, A.indptr) and np.array_equal(idx, A.indices) and np.array_equal(dat, A.data)
    # duplicates: summed (in input order)
    rows = np.array([0, 2, 0, 1, 0, 2], np.int32)
    cols = np.array([1, 2, 1, 1, 0, 2], np.int32)
    vals = np.array([1.0, 2.0, 0.5, 3.0, 4.0, 1e-17])
    G = sp.GpuCsrMat.from_triplets(3, rows, cols, vals)
    ip, idx, dat = G.download()
    assert ip.tolist() == [0, 2, 3, 4] and idx.tolist() == [0, 1, 1, 2]
    assert dat.tolist() == [4.0, 1.0 + 0.5, 3.0, 2.0 + 1e-17]
    with pytest.raises(Exception):
        sp.GpuCsrMat.from_triplets(3, [0, 3], [0, 0], [1.0, 1.0])  # row index out of range


@pytest.mark.parametrize("sym", ["general", "symmetric", "hermitian"])
def test_matrix_market_roundtrip(sp, orc, tmp_path, sym):
    rng = np.random.default_rng(5)
    n = 60
    M = sps.random(n, n, density=0.08, random_state=2, format="coo", dtype=np.float64)
    if sym == "general":
        M = (M + sps.identity(n) * 3.0).tocoo()
        dtype = np.float64
    elif sym == "symmetric":
        M = (M + M.T + sps.identity(n) * 3.0).tocoo()
        dtype = np.float64
    else:
        Z = M + 1j * sps.random(n, n, density=0.08, random_state=4, format="coo")
        M = (Z + Z.getH() + sps.identity(n) * 3.0).tocoo()
        dtype = np.complex128
    path = os.path.join(tmp_path, f"m_{sym}.mtx")
    scipy.io.mmwrite(path, M, symmetry=sym, precision=17)
    G = sp.GpuCsrMat.read_matrix_market(path, dtype=dtype)
    ref = M.tocsr()
    ref.sort_indices()
    ip, idx, dat = G.download()
    assert np.array_equal(ip, ref.indptr) and np.array_equal(idx, ref.indices)
    assert np.allclose(dat, ref.data, rtol=1e-15, atol=0)
    x = rng.uniform(-1, 1, n).astype(dtype)
    y = np.zeros(n, dtype)
    G.mul_vec(x, y)
    assert np.allclose(y, ref @ x, rtol=1e-13, atol=1e-13)
    with open(os.path.join(tmp_path, "bad.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(sp.IncompatibleMatrixFormat):
        sp.GpuCsrMat.read_matrix_market(os.path.join(tmp_path, "bad.mtx"))
