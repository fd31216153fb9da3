// create.cu -- building device-resident CSR operators: upload of a host CSR matrix (the
// MklMat::new analogue, src/mkl_mat.rs:32-74) and on-device synthetic stencil generators
// (the reference's generator is src/main.rs:53-88).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "csr.cuh"
#include "dist.cuh"

namespace spb {

static const size_t kPad = 8;  // staging loads may read up to 3 elements past nnz

// ---------------------------------------------------------------- stencil row descriptions
struct StencilDesc {
  int kind;
  int64_t nx, ny, nz;
  double p0, p1, p2;
};

__device__ __forceinline__ bool dirichlet_border(int64_t i, int64_t j, int64_t rows, int64_t cols) {
  return i == 0 || i + 1 == rows || j == 0 || j + 1 == cols;  // src/main.rs:40-51
}

__device__ __forceinline__ int stencil_row_count(const StencilDesc& d, int64_t row) {
  if (d.kind == SPB_STENCIL_DIRICHLET2D) {
    const int64_t i = row / d.ny, j = row % d.ny;
    return dirichlet_border(i, j, d.nx, d.ny) ? 1 : 5;
  }
  const int64_t x = row % d.nx, y = (row / d.nx) % d.ny, z = row / (d.nx * d.ny);
  const int cx = 1 + (x > 0) + (x + 1 < d.nx), cy = 1 + (y > 0) + (y + 1 < d.ny),
            cz = 1 + (z > 0) + (z + 1 < d.nz);
  if (d.kind == SPB_STENCIL_LAP3D7) return cx + cy + cz - 2;
  return cx * cy * cz;  // SPB_STENCIL_CONVDIFF27
}

__global__ void stencil_count_kernel(StencilDesc d, int64_t row_begin, int64_t n_local,
                                     int64_t* counts) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n_local) counts[r] = stencil_row_count(d, row_begin + r);
  if (r == n_local) counts[r] = 0;
}

template <typename T>
__device__ __forceinline__ T make_val(double re, double im);
template <>
__device__ __forceinline__ double make_val<double>(double re, double) {
  return re;
}
template <>
__device__ __forceinline__ cplx make_val<cplx>(double re, double im) {
  return cplx{re, im};
}
template <>
__device__ __forceinline__ float make_val<float>(double re, double) {
  return (float)re;
}
template <>
__device__ __forceinline__ cplxf make_val<cplxf>(double re, double im) {
  return cplxf{(float)re, (float)im};
}

template <typename T>
__global__ void stencil_fill_kernel(StencilDesc d, int64_t row_begin, int64_t n_local,
                                    const int64_t* indptr, int* cols, T* vals) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_local) return;
  const int64_t row = row_begin + r;
  int64_t k = indptr[r];
  if (d.kind == SPB_STENCIL_DIRICHLET2D) {
    // src/main.rs:53-88: border rows = identity, interior [1, 1, -4, 1, 1]; the reference
    // computes column ids as i*rows + j (square grids), mirrored here.
    const int64_t i = row / d.ny, j = row % d.ny;
    if (dirichlet_border(i, j, d.nx, d.ny)) {
      cols[k] = (int)(i * d.nx + j);
      vals[k] = make_val<T>(1.0, 0.0);
    } else {
      cols[k] = (int)((i - 1) * d.nx + j); vals[k++] = make_val<T>(1.0, 0.0);
      cols[k] = (int)(i * d.nx + j - 1);   vals[k++] = make_val<T>(1.0, 0.0);
      cols[k] = (int)(i * d.nx + j);       vals[k++] = make_val<T>(-4.0, 0.0);
      cols[k] = (int)(i * d.nx + j + 1);   vals[k++] = make_val<T>(1.0, 0.0);
      cols[k] = (int)((i + 1) * d.nx + j); vals[k++] = make_val<T>(1.0, 0.0);
    }
    return;
  }
  const int64_t x = row % d.nx, y = (row / d.nx) % d.ny, z = row / (d.nx * d.ny);
  const int64_t sxy = d.nx * d.ny;
  if (d.kind == SPB_STENCIL_LAP3D7) {
    const T off = make_val<T>(-1.0, 0.0);
    if (z > 0) { cols[k] = (int)(row - sxy); vals[k++] = off; }
    if (y > 0) { cols[k] = (int)(row - d.nx); vals[k++] = off; }
    if (x > 0) { cols[k] = (int)(row - 1); vals[k++] = off; }
    cols[k] = (int)row; vals[k++] = make_val<T>(6.0 - d.p0, 0.0 - d.p1);
    if (x + 1 < d.nx) { cols[k] = (int)(row + 1); vals[k++] = off; }
    if (y + 1 < d.ny) { cols[k] = (int)(row + d.nx); vals[k++] = off; }
    if (z + 1 < d.nz) { cols[k] = (int)(row + sxy); vals[k++] = off; }
    return;
  }
  // 27-point convection-diffusion: (27 I - S27) + bx Dx + by Dy + bz Dz, upwind bidiagonals
  const double centre = 26.0 + d.p0 + d.p1 + d.p2;
  for (int dz = -1; dz <= 1; ++dz) {
    if (z + dz < 0 || z + dz >= d.nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      if (y + dy < 0 || y + dy >= d.ny) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        if (x + dx < 0 || x + dx >= d.nx) continue;
        double v;
        if (dx == 0 && dy == 0 && dz == 0) v = centre;
        else if (dx == -1 && dy == 0 && dz == 0) v = -1.0 - d.p0;
        else if (dx == 0 && dy == -1 && dz == 0) v = -1.0 - d.p1;
        else if (dx == 0 && dy == 0 && dz == -1) v = -1.0 - d.p2;
        else v = -1.0;
        cols[k] = (int)(row + dx + dy * d.nx + dz * sxy);
        vals[k++] = make_val<T>(v, 0.0);
      }
    }
  }
}

__global__ void narrow_indptr_kernel(const int64_t* in, int64_t n1, int* out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n1) out[i] = (int)in[i];
}

// Structural validation of caller-supplied index arrays (sprs::CsMat::new rejects such input before
// any solver of the reference sees it): every column id must lie in [0, ncols).
__global__ void validate_cols_kernel(const int* cols, int64_t nnz, int ncols, int* bad) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
    if ((unsigned)cols[k] >= (unsigned)ncols) atomicExch(bad, 1);  // integer flag: order-independent
}

void validate_device_cols(Ctx* ctx, const int* cols, int64_t nnz, int64_t ncols) {
  if (nnz <= 0) return;
  DevBuf bad;
  bad.alloc(sizeof(int));
  SPB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), ctx->stream));
  {
    LaunchScope ls(ctx, FAM_SCALAR);
    const int grid = (int)std::min<int64_t>(ceil_div(nnz, 256), (int64_t)ctx->sm_count * 16);
    validate_cols_kernel<<<grid, 256, 0, ctx->stream>>>(cols, nnz, (int)ncols, bufptr<int>(bad));
    check_launch("validate_cols_kernel");
  }
  int h = 0;
  SPB_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SPB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "column index out of range");
}

// indptr[0] == 0, non-decreasing (host array of n1 entries)
template <typename IP>
static void validate_host_indptr(const IP* ip, int64_t n1) {
  if (n1 < 1 || ip[0] != 0) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "indptr must start at 0");
  for (int64_t i = 1; i < n1; ++i)
    if (ip[i] < ip[i - 1]) SPB_FAIL(SPB_INCOMPATIBLE_FORMAT, "indptr must be non-decreasing");
}

template <typename T>
static void finish_create(CsrMat<T>* m) {
  if (m->ctx->dist)
    csr_localize(m);
  m->analyze();
}

// ---------------------------------------------------------------- from host arrays
template <typename T>
CsrMat<T>* csr_from_host(Ctx* ctx, int64_t n_global, int64_t row_begin, int64_t row_end,
                         const void* indptr, int indptr_bits, const int32_t* indices,
                         const void* values) {
  if (indptr_bits != 32 && indptr_bits != 64) SPB_FAIL(SPB_INVALID_ARG, "indptr_bits must be 32 or 64");
  if (row_begin < 0 || row_end < row_begin || row_end > n_global)
    SPB_FAIL(SPB_INVALID_ARG, "bad row range");
  if (n_global >= (int64_t)1 << 31) SPB_FAIL(SPB_INVALID_ARG, "matrix dimension exceeds int32 columns");
  const int64_t nl = row_end - row_begin;
  if (indptr_bits == 64)
    validate_host_indptr((const int64_t*)indptr, nl + 1);
  else
    validate_host_indptr((const int32_t*)indptr, nl + 1);
  const int64_t nnz = indptr_bits == 64 ? ((const int64_t*)indptr)[nl] : (int64_t)((const int32_t*)indptr)[nl];
  if (nnz > 0 && (!indices || !values)) SPB_FAIL(SPB_INVALID_ARG, "null index / value array");
  auto* m = new CsrMat<T>();
  try {
    m->ctx = ctx;
    m->kind = OP_CSR;
    m->dtype = ScalarTraits<T>::dtype;
    m->n_global = n_global;
    m->n_local = nl;
    m->row_begin = row_begin;
    m->nnz = nnz;
    m->ip64 = nnz >= ((int64_t)1 << 31) - 8;
    if (m->ip64) {
      std::vector<int64_t> ip(nl + 1);
      for (int64_t i = 0; i <= nl; ++i)
        ip[i] = indptr_bits == 64 ? ((const int64_t*)indptr)[i] : (int64_t)((const int32_t*)indptr)[i];
      m->indptr.alloc(sizeof(int64_t) * (nl + 1));
      SPB_CUDA(cudaMemcpyAsync(m->indptr.p, ip.data(), sizeof(int64_t) * (nl + 1), cudaMemcpyHostToDevice, ctx->stream));
      SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
      std::vector<int32_t> ip(nl + 1);
      for (int64_t i = 0; i <= nl; ++i)
        ip[i] = indptr_bits == 64 ? (int32_t)((const int64_t*)indptr)[i] : ((const int32_t*)indptr)[i];
      m->indptr.alloc(sizeof(int32_t) * (nl + 1));
      SPB_CUDA(cudaMemcpyAsync(m->indptr.p, ip.data(), sizeof(int32_t) * (nl + 1), cudaMemcpyHostToDevice, ctx->stream));
      SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    m->cols.alloc(sizeof(int) * (nnz + kPad));
    m->vals.alloc(sizeof(T) * (nnz + kPad));
    SPB_CUDA(cudaMemsetAsync(m->cols.p, 0, m->cols.bytes, ctx->stream));
    SPB_CUDA(cudaMemsetAsync(m->vals.p, 0, m->vals.bytes, ctx->stream));
    if (nnz > 0) {
      SPB_CUDA(cudaMemcpyAsync(m->cols.p, indices, sizeof(int) * nnz, cudaMemcpyHostToDevice, ctx->stream));
      SPB_CUDA(cudaMemcpyAsync(m->vals.p, values, sizeof(T) * nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    validate_device_cols(ctx, bufptr<int>(m->cols), nnz, n_global);  // before localisation / analysis index with them
    finish_create(m);
  } catch (...) {
    delete m;
    throw;
  }
  return m;
}

// ---------------------------------------------------------------- from device arrays (ingest.cu)
template <typename T>
CsrMat<T>* csr_adopt_device(Ctx* ctx, int64_t n, int64_t nnz, DevBuf&& indptr32, DevBuf&& cols, DevBuf&& vals) {
  auto* m = new CsrMat<T>();
  try {
    m->ctx = ctx;
    m->kind = OP_CSR;
    m->dtype = ScalarTraits<T>::dtype;
    m->n_global = m->n_local = n;
    m->row_begin = 0;
    m->nnz = nnz;
    m->ip64 = false;
    m->indptr = std::move(indptr32);
    m->cols = std::move(cols);  // both already carry the kPad tail
    m->vals = std::move(vals);
    finish_create(m);
  } catch (...) {
    delete m;
    throw;
  }
  return m;
}
template CsrMat<double>* csr_adopt_device<double>(Ctx*, int64_t, int64_t, DevBuf&&, DevBuf&&, DevBuf&&);
template CsrMat<cplx>* csr_adopt_device<cplx>(Ctx*, int64_t, int64_t, DevBuf&&, DevBuf&&, DevBuf&&);
template CsrMat<float>* csr_adopt_device<float>(Ctx*, int64_t, int64_t, DevBuf&&, DevBuf&&, DevBuf&&);
template CsrMat<cplxf>* csr_adopt_device<cplxf>(Ctx*, int64_t, int64_t, DevBuf&&, DevBuf&&, DevBuf&&);

// ---------------------------------------------------------------- on-device generators
void stencil_partition(int kind, int64_t nx, int64_t ny, int64_t nz, int world, int rank,
                       int64_t* row_begin, int64_t* row_end) {
  // contiguous row blocks cut at plane boundaries (z-slabs for the 3-D stencils)
  const int64_t plane = kind == SPB_STENCIL_DIRICHLET2D ? ny : nx * ny;
  const int64_t nplanes = kind == SPB_STENCIL_DIRICHLET2D ? nx : nz;
  *row_begin = plane * ((nplanes * rank) / world);
  *row_end = plane * ((nplanes * (rank + 1)) / world);
}

template <typename T>
CsrMat<T>* csr_from_stencil(Ctx* ctx, int kind, int64_t nx, int64_t ny, int64_t nz,
                            const double* params, int nparams) {
  if (kind < 0 || kind > SPB_STENCIL_CONVDIFF27) SPB_FAIL(SPB_INVALID_ARG, "unknown stencil kind");
  if (kind == SPB_STENCIL_DIRICHLET2D) {
    nz = 1;
    if (nx != ny) SPB_FAIL(SPB_INVALID_ARG, "reference Dirichlet generator needs a square grid");
  }
  if (nx < 1 || ny < 1 || nz < 1) SPB_FAIL(SPB_INVALID_ARG, "bad grid");
  const int64_t n = nx * ny * nz;
  if (n >= ((int64_t)1 << 31)) SPB_FAIL(SPB_INVALID_ARG, "matrix dimension exceeds int32 columns");
  StencilDesc d{kind, nx, ny, nz, nparams > 0 ? params[0] : 0.0, nparams > 1 ? params[1] : 0.0,
                nparams > 2 ? params[2] : 0.0};
  if (kind == SPB_STENCIL_LAP3D7 && !ScalarTraits<T>::is_complex && d.p1 != 0.0)
    SPB_FAIL(SPB_INVALID_ARG, "complex shift needs dtype SPB_C128");
  int64_t rb = 0, re = n;
  if (ctx->dist) stencil_partition(kind, nx, ny, nz, ctx->world(), ctx->rank(), &rb, &re);
  const int64_t nl = re - rb;
  auto* m = new CsrMat<T>();
  try {
    m->ctx = ctx;
    m->kind = OP_CSR;
    m->dtype = ScalarTraits<T>::dtype;
    m->n_global = n;
    m->n_local = nl;
    m->row_begin = rb;
    DevBuf counts, ip64buf;
    counts.alloc(sizeof(int64_t) * (nl + 1));
    ip64buf.alloc(sizeof(int64_t) * (nl + 1));
    const int grid = (int)ceil_div(nl + 1, 256);
    {
      LaunchScope ls(ctx, FAM_SCALAR);
      stencil_count_kernel<<<grid, 256, 0, ctx->stream>>>(d, rb, nl, bufptr<int64_t>(counts));
      check_launch("stencil_count_kernel");
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, bufptr<int64_t>(counts), bufptr<int64_t>(ip64buf), nl + 1, ctx->stream);
    DevBuf tmp;
    tmp.alloc(tmp_bytes);
    cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, bufptr<int64_t>(counts), bufptr<int64_t>(ip64buf), nl + 1, ctx->stream);
    int64_t nnz = 0;
    SPB_CUDA(cudaMemcpyAsync(&nnz, bufptr<int64_t>(ip64buf) + nl, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    counts.release();
    tmp.release();
    m->nnz = nnz;
    m->cols.alloc(sizeof(int) * (nnz + kPad));
    m->vals.alloc(sizeof(T) * (nnz + kPad));
    SPB_CUDA(cudaMemsetAsync((char*)m->cols.p + sizeof(int) * nnz, 0, sizeof(int) * kPad, ctx->stream));
    SPB_CUDA(cudaMemsetAsync((char*)m->vals.p + sizeof(T) * nnz, 0, sizeof(T) * kPad, ctx->stream));
    if (nl > 0) {
      LaunchScope ls(ctx, FAM_SCALAR);
      stencil_fill_kernel<T><<<(int)ceil_div(nl, 128), 128, 0, ctx->stream>>>(
          d, rb, nl, bufptr<int64_t>(ip64buf), bufptr<int>(m->cols), bufptr<T>(m->vals));
      check_launch("stencil_fill_kernel");
    }
    m->ip64 = nnz >= ((int64_t)1 << 31) - 8;
    if (m->ip64) {
      m->indptr = std::move(ip64buf);
    } else {
      m->indptr.alloc(sizeof(int) * (nl + 1));
      LaunchScope ls(ctx, FAM_SCALAR);
      narrow_indptr_kernel<<<grid, 256, 0, ctx->stream>>>(bufptr<int64_t>(ip64buf), nl + 1, bufptr<int>(m->indptr));
      check_launch("narrow_indptr_kernel");
    }
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
    finish_create(m);
  } catch (...) {
    delete m;
    throw;
  }
  return m;
}

#define SPB_INST_CREATE(T)                                                                                                 \
  template CsrMat<T>* csr_from_host<T>(Ctx*, int64_t, int64_t, int64_t, const void*, int, const int32_t*, const void*);   \
  template CsrMat<T>* csr_from_stencil<T>(Ctx*, int, int64_t, int64_t, int64_t, const double*, int);
SPB_INST_CREATE(double)
SPB_INST_CREATE(cplx)
SPB_INST_CREATE(float)
SPB_INST_CREATE(cplxf)

}  // namespace spb
