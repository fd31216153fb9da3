// ingest.cu -- the callers' side of the operator boundary: building the
// device CSR operator from the other layouts the reference and its tests start from.
//   * triplets  -> CSR: sprs::TriMat::to_csr as used by the reference's fixtures
//                  (tests/test_minres.rs:65-119, tests/test_complex_solve.rs:99-213): entries sorted
//                  by (row, column), duplicates summed (here: in input order);
//   * CSC       -> CSR: the reference multiplies a CSC matrix column by column,
//                  `v_out[row] += v_in[col] * value` (src/mat.rs:130-142, KAT :208-229).  Each
//                  output element therefore accumulates its row's entries in increasing column
//                  order starting from zero -- exactly the CSR row fold of the stable transpose, so
//                  the operator is bit-identical to the reference loop and runs at CSR speed;
//   * Matrix Market coordinate files (the format the reference's fixtures were exported from,
//                  tests/test_complex_solve.rs:14): parsed on the host, assembled on the device.
// The sort is cub::DeviceRadixSort (stable) on (row << 32 | col) keys; everything after the upload
// stays on the device.  Single-GPU matrices only (a partitioned matrix is uploaded per row block
// with spb_csr_create).
#include <cub/cub.cuh>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "csr.cuh"
#include "vecops.cuh"

namespace spb {

static const size_t kIngestPad = 8;  // same tail padding as create.cu

template <typename T>
CsrMat<T>* csr_adopt_device(Ctx* ctx, int64_t n, int64_t nnz, DevBuf&& indptr32, DevBuf&& cols, DevBuf&& vals);  // create.cu

__global__ void ingest_keys_kernel(int64_t nnz, const int* rows, const int* cols, unsigned long long* keys, int* idx, int64_t n,
                                   int* bad) {
  SPB_GRID_STRIDE(k, nnz) {
    const int r = rows[k], c = cols[k];
    if (r < 0 || r >= n || c < 0 || c >= n) atomicExch(bad, 1);
    keys[k] = ((unsigned long long)(unsigned)r << 32) | (unsigned)c;
    idx[k] = (int)k;
  }
}

// column index of every CSC entry (one thread per column)
template <typename IP>
__global__ void csc_expand_kernel(int64_t ncols, const IP* indptr, int* col_of) {
  SPB_GRID_STRIDE(c, ncols) {
    for (IP k = indptr[c]; k < indptr[c + 1]; ++k) col_of[k] = (int)c;
  }
}

// head[k] = 1 where a new (row, col) run starts (sum_duplicates) or everywhere (keep all)
__global__ void ingest_heads_kernel(int64_t nnz, const unsigned long long* keys, int sum_duplicates, int* head) {
  SPB_GRID_STRIDE(k, nnz) head[k] = (!sum_duplicates || k == 0 || keys[k] != keys[k - 1]) ? 1 : 0;
}

template <typename T>
__global__ void ingest_fill_kernel(int64_t nnz, const unsigned long long* keys, const int* perm, const int* head, const int* pos,
                                   const T* vals_in, int* rows_out, int* cols_out, T* vals_out) {
  SPB_GRID_STRIDE(k, nnz) {
    if (!head[k]) continue;
    T s = vals_in[perm[k]];
    for (int64_t j = k + 1; j < nnz && !head[j]; ++j) s = add(s, vals_in[perm[j]]);  // duplicates: input order (stable sort)
    const int o = pos[k];
    rows_out[o] = (int)(keys[k] >> 32);
    cols_out[o] = (int)(keys[k] & 0xffffffffu);
    vals_out[o] = s;
  }
}

// indptr[r] = first entry whose row is >= r (rows_out is sorted)
__global__ void ingest_indptr_kernel(int64_t n, int64_t nnz_out, const int* rows_out, int* indptr) {
  SPB_GRID_STRIDE(r, n + 1) {
    int64_t lo = 0, hi = nnz_out;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (rows_out[mid] < r) lo = mid + 1; else hi = mid;
    }
    indptr[r] = (int)lo;
  }
}

// rows / cols / vals: DEVICE arrays of nnz triplets.  by_row_only: stable sort on the row alone
// (columns keep their input order -- the CSC transpose), else on (row, col).
template <typename T>
static CsrMat<T>* assemble(Ctx* c, int64_t n, int64_t nnz, const int* d_rows, const int* d_cols, const T* d_vals, bool by_row_only,
                           bool sum_duplicates) {
  if (c->dist) SPB_FAIL(SPB_INVALID_ARG, "triplet / CSC / Matrix Market ingestion builds single-GPU matrices");
  if (n < 0 || n >= ((int64_t)1 << 31) - 1 || nnz < 0 || nnz >= ((int64_t)1 << 31) - 16)
    SPB_FAIL(SPB_INVALID_ARG, "matrix too large for the int32 ingestion path");
  const size_t m = (size_t)std::max<int64_t>(nnz, 1);
  DevBuf keys, keys2, idx, idx2, head, pos, rows_out, bad, tmp;
  keys.alloc(8 * m);
  keys2.alloc(8 * m);
  idx.alloc(4 * m);
  idx2.alloc(4 * m);
  head.alloc(4 * m);
  pos.alloc(4 * (m + 1));
  rows_out.alloc(4 * m);
  bad.alloc(16);
  SPB_CUDA(cudaMemsetAsync(bad.p, 0, 16, c->stream));
  DevBuf cols_out, vals_out, indptr;
  cols_out.alloc(sizeof(int) * (m + kIngestPad));
  vals_out.alloc(sizeof(T) * (m + kIngestPad));
  indptr.alloc(sizeof(int) * (size_t)(n + 1));
  SPB_CUDA(cudaMemsetAsync(cols_out.p, 0, cols_out.bytes, c->stream));
  SPB_CUDA(cudaMemsetAsync(vals_out.p, 0, vals_out.bytes, c->stream));
  int64_t nnz_out = 0;
  if (nnz > 0) {
    {
      LaunchScope ls(c, FAM_PACK);
      ingest_keys_kernel<<<vec_grid(c, nnz), kVecThreads, 0, c->stream>>>(nnz, d_rows, d_cols, keys.as<unsigned long long>(), idx.as<int>(), n,
                                                                           bad.as<int>());
      check_launch("ingest_keys_kernel");
    }
    const int begin_bit = by_row_only ? 32 : 0;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<int>(), idx2.as<int>(),
                                    (int)nnz, begin_bit, 64, c->stream);
    tmp.alloc(tb);
    {
      LaunchScope ls(c, FAM_PACK);
      SPB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), idx.as<int>(),
                                               idx2.as<int>(), (int)nnz, begin_bit, 64, c->stream));
    }
    {
      LaunchScope ls(c, FAM_PACK);
      ingest_heads_kernel<<<vec_grid(c, nnz), kVecThreads, 0, c->stream>>>(nnz, keys2.as<unsigned long long>(), sum_duplicates ? 1 : 0,
                                                                            head.as<int>());
      check_launch("ingest_heads_kernel");
    }
    size_t tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, head.as<int>(), pos.as<int>(), (int)nnz, c->stream);
    tmp.ensure(tb2);
    {
      LaunchScope ls(c, FAM_PACK);
      SPB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, head.as<int>(), pos.as<int>(), (int)nnz, c->stream));
    }
    int last_pos = 0, last_head = 0, isbad = 0;
    SPB_CUDA(cudaMemcpyAsync(&last_pos, pos.as<int>() + (nnz - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaMemcpyAsync(&last_head, head.as<int>() + (nnz - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaMemcpyAsync(&isbad, bad.p, 4, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    if (isbad) SPB_FAIL(SPB_INVALID_ARG, "row or column index out of range");
    nnz_out = (int64_t)last_pos + last_head;
    {
      LaunchScope ls(c, FAM_PACK);
      ingest_fill_kernel<T><<<vec_grid(c, nnz), kVecThreads, 0, c->stream>>>(nnz, keys2.as<unsigned long long>(), idx2.as<int>(), head.as<int>(),
                                                                              pos.as<int>(), d_vals, rows_out.as<int>(), cols_out.as<int>(),
                                                                              vals_out.as<T>());
      check_launch("ingest_fill_kernel");
    }
  }
  {
    LaunchScope ls(c, FAM_PACK);
    ingest_indptr_kernel<<<vec_grid(c, n + 1), kVecThreads, 0, c->stream>>>(n, nnz_out, rows_out.as<int>(), indptr.as<int>());
    check_launch("ingest_indptr_kernel");
  }
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  return csr_adopt_device<T>(c, n, nnz_out, std::move(indptr), std::move(cols_out), std::move(vals_out));
}

template <typename T>
CsrMat<T>* csr_from_triplets(Ctx* c, int64_t n, int64_t nnz, const int32_t* rows, const int32_t* cols, const void* vals) {
  const size_t m = (size_t)std::max<int64_t>(nnz, 1);
  DevBuf dr, dc, dv;
  dr.alloc(4 * m);
  dc.alloc(4 * m);
  dv.alloc(sizeof(T) * m);
  if (nnz > 0) {
    SPB_CUDA(cudaMemcpyAsync(dr.p, rows, 4 * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(dc.p, cols, 4 * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(dv.p, vals, sizeof(T) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  return assemble<T>(c, n, nnz, dr.as<int>(), dc.as<int>(), dv.as<T>(), false, true);
}

template <typename T>
CsrMat<T>* csr_from_csc(Ctx* c, int64_t n, const void* indptr, int indptr_bits, const int32_t* row_indices, const void* vals) {
  if (indptr_bits != 32 && indptr_bits != 64) SPB_FAIL(SPB_INVALID_ARG, "indptr_bits must be 32 or 64");
  // the column pointers index the row-index / value arrays: indptr[0] == 0 and non-decreasing, else the
  // expansion kernel writes out of bounds (the row indices themselves are range-checked by assemble())
  for (int64_t i = 0; i <= n; ++i) {
    const int64_t v = indptr_bits == 64 ? ((const int64_t*)indptr)[i] : (int64_t)((const int32_t*)indptr)[i];
    const int64_t prev = i == 0 ? 0 : (indptr_bits == 64 ? ((const int64_t*)indptr)[i - 1] : (int64_t)((const int32_t*)indptr)[i - 1]);
    if ((i == 0 && v != 0) || v < prev) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "CSC indptr must start at 0 and be non-decreasing");
  }
  const int64_t nnz = indptr_bits == 64 ? ((const int64_t*)indptr)[n] : (int64_t)((const int32_t*)indptr)[n];
  if (nnz < 0 || nnz >= ((int64_t)1 << 31) - 16) SPB_FAIL(SPB_INVALID_ARG, "matrix too large for the int32 ingestion path");
  const size_t m = (size_t)std::max<int64_t>(nnz, 1);
  DevBuf dip, dr, dc, dv;
  dip.alloc((size_t)(indptr_bits / 8) * (size_t)(n + 1));
  dr.alloc(4 * m);
  dc.alloc(4 * m);
  dv.alloc(sizeof(T) * m);
  SPB_CUDA(cudaMemcpyAsync(dip.p, indptr, (size_t)(indptr_bits / 8) * (size_t)(n + 1), cudaMemcpyHostToDevice, c->stream));
  if (nnz > 0) {
    SPB_CUDA(cudaMemcpyAsync(dr.p, row_indices, 4 * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(dv.p, vals, sizeof(T) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    LaunchScope ls(c, FAM_PACK);
    if (indptr_bits == 64)
      csc_expand_kernel<int64_t><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(n, dip.as<int64_t>(), dc.as<int>());
    else
      csc_expand_kernel<int32_t><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(n, dip.as<int32_t>(), dc.as<int>());
    check_launch("csc_expand_kernel");
  }
  // stable by row: inside a row the entries keep increasing-column order, duplicates are kept
  return assemble<T>(c, n, nnz, dr.as<int>(), dc.as<int>(), dv.as<T>(), true, false);
}

// ---------------------------------------------------------------- Matrix Market (coordinate)
static std::string lower(std::string s) {
  for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
  return s;
}

template <typename T>
CsrMat<T>* csr_from_matrix_market(Ctx* c, const char* path) {
  FILE* f = fopen(path, "r");
  if (!f) SPB_FAIL(SPB_INVALID_ARG, std::string("cannot open ") + path);
  std::vector<int32_t> rows, cols;
  std::vector<T> vals;
  int64_t n = 0;
  try {
    char line[1024];
    if (!fgets(line, sizeof(line), f)) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "empty Matrix Market file");
    char banner[64], object[64], format[64], field[64], symmetry[64];
    if (sscanf(line, "%63s %63s %63s %63s %63s", banner, object, format, field, symmetry) != 5 || lower(banner) != "%%matrixmarket" ||
        lower(object) != "matrix")
      SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "not a Matrix Market matrix header");
    const std::string fmt = lower(format), fld = lower(field), sym = lower(symmetry);
    if (fmt != "coordinate") SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "only the coordinate (sparse) Matrix Market format is supported");
    const bool is_complex = fld == "complex", is_pattern = fld == "pattern";
    if (!is_complex && !is_pattern && fld != "real" && fld != "integer" && fld != "double")
      SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "unknown Matrix Market field " + fld);
    if (is_complex && !ScalarTraits<T>::is_complex) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "complex Matrix Market file requested as f64");
    const bool symm = sym == "symmetric", skew = sym == "skew-symmetric", herm = sym == "hermitian";
    if (!symm && !skew && !herm && sym != "general") SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "unknown Matrix Market symmetry " + sym);
    do {
      if (!fgets(line, sizeof(line), f)) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "missing size line");
    } while (line[0] == '%' || line[0] == '\n' || line[0] == '\r');
    long long M = 0, N = 0, NNZ = 0;
    if (sscanf(line, "%lld %lld %lld", &M, &N, &NNZ) != 3) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "bad size line");
    if (M != N) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "Not a square matrix");  // src/mkl_mat.rs:37
    if (M >= ((long long)1 << 31) - 1) SPB_FAIL(SPB_INVALID_ARG, "matrix too large");
    n = M;
    rows.reserve((size_t)NNZ * ((symm || skew || herm) ? 2 : 1));
    cols.reserve(rows.capacity());
    vals.reserve(rows.capacity());
    for (long long k = 0; k < NNZ; ++k) {
      long long i = 0, j = 0;
      double re = 1.0, im = 0.0;
      int got;
      if (is_pattern)
        got = fscanf(f, "%lld %lld", &i, &j) == 2 ? 3 : 0;
      else if (is_complex)
        got = fscanf(f, "%lld %lld %lf %lf", &i, &j, &re, &im) == 4 ? 3 : 0;
      else
        got = fscanf(f, "%lld %lld %lf", &i, &j, &re) == 3 ? 3 : 0;
      if (got != 3 || i < 1 || j < 1 || i > M || j > N) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "bad Matrix Market entry");
      const scal2 s{re, im};
      rows.push_back((int32_t)(i - 1));
      cols.push_back((int32_t)(j - 1));
      vals.push_back(from_scal2<T>(s));
      if (i != j && (symm || skew || herm)) {
        rows.push_back((int32_t)(j - 1));
        cols.push_back((int32_t)(i - 1));
        vals.push_back(from_scal2<T>(skew ? scal2{-re, -im} : herm ? scal2{re, -im} : s));
      }
    }
  } catch (...) {
    fclose(f);
    throw;
  }
  fclose(f);
  return csr_from_triplets<T>(c, n, (int64_t)rows.size(), rows.data(), cols.data(), vals.data());
}

#define SPB_INST_INGEST(T)                                                                                           \
  template CsrMat<T>* csr_from_triplets<T>(Ctx*, int64_t, int64_t, const int32_t*, const int32_t*, const void*);     \
  template CsrMat<T>* csr_from_csc<T>(Ctx*, int64_t, const void*, int, const int32_t*, const void*);                 \
  template CsrMat<T>* csr_from_matrix_market<T>(Ctx*, const char*);
SPB_INST_INGEST(double)
SPB_INST_INGEST(cplx)
SPB_INST_INGEST(float)
SPB_INST_INGEST(cplxf)

}  // namespace spb
