// spmv.cu -- CSR sparse matrix-vector product for sm_100a (f64 and Complex<f64>).
//
// Replaces MatVecMul::mul_vec_unchecked / mul_vec_dot_unchecked for CSR (src/mat.rs:68-129,
// :145-152) and MklMat's mkl_sparse_?_mv / mkl_sparse_?_dotmv (src/mkl_mat.rs:170-319).
//
// Design (HBM-bound, no tensor cores -- SpMV is not a dense contraction):
//   * analysis once per matrix (the mkl_sparse_optimize analogue): rows are cut into tiles of
//     ~TILE non-zeros by a binary search on indptr, so every CTA streams the same number of bytes
//     whatever the row lengths are;
//   * a persistent grid (a multiple of the SM count) walks the tiles round-robin, so concurrently
//     running CTAs work on neighbouring rows and the x window they gather from stays in L2;
//   * per tile, col_idx / values are streamed with coalesced 128-bit ld.global.nc.L1::no_allocate
//     loads into shared memory (the matrix is read exactly once and never pollutes L1);
//   * then one thread per row accumulates  acc = acc + x[col] * val  SEQUENTIALLY IN CSR ORDER --
//     the same left fold as src/mat.rs:100-105 -- so y is bit-identical to the reference; lanes of
//     a warp walk the same stencil diagonal, so the x gathers (ld.global.nc through L1) coalesce;
//   * dot-product epilogues (<r0,v>, <t,t>, <t,r>, conj(x).y) are folded in, so the Krylov loops
//     never re-read the SpMV output for a reduction.
// Algorithmic bytes per launch: nnz*(sizeof(T)+4) + (n+1)*sizeof(indptr) + 2*n*sizeof(T).
#include <cub/cub.cuh>

#include "csr.cuh"
#include "reduce.cuh"

namespace spb {

// ---------------------------------------------------------------- streaming loads
__device__ __forceinline__ int4 ld_stream_int4(const int* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream_double2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_x(const double* p) { return __ldg(p); }
__device__ __forceinline__ cplx ld_x(const cplx* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}

template <typename T>
struct ValPack;  // 4 consecutive values staged by one thread
template <>
struct ValPack<double> {
  double2 a, b;
  __device__ __forceinline__ void load(const double* p) {
    a = ld_stream_double2(p);
    b = ld_stream_double2(p + 2);
  }
  __device__ __forceinline__ void store(double* s) const {
    *reinterpret_cast<double2*>(s) = a;
    *reinterpret_cast<double2*>(s + 2) = b;
  }
};
template <>
struct ValPack<cplx> {
  double2 a, b, c, d;
  __device__ __forceinline__ void load(const cplx* p) {
    const double* q = reinterpret_cast<const double*>(p);
    a = ld_stream_double2(q);
    b = ld_stream_double2(q + 2);
    c = ld_stream_double2(q + 4);
    d = ld_stream_double2(q + 6);
  }
  __device__ __forceinline__ void store(cplx* s) const {
    double2* q = reinterpret_cast<double2*>(s);
    q[0] = a;
    q[1] = b;
    q[2] = c;
    q[3] = d;
  }
};

template <typename T, typename IP>
struct SpmvArgs {
  const IP* indptr;
  const int* cols;
  const T* vals;
  const int* tile_row;
  const int* tile_list;  // optional indirection (interior / boundary passes)
  int64_t ntiles;
  const T* x;    // owned entries  [0, n_local)
  const T* xh;   // halo entries   [n_local, n_local + n_halo)
  int n_local;
  T* y;
  const T* w;    // epilogue operand
  T* partials;   // [2 * gridDim.x]
  const int* gate;  // optional solver gate (see Ctx::gate)
  int gate_value;
};

template <typename T, bool CONJ_IN>
__device__ __forceinline__ T gather_x(const T* x, const T* xh, int n_local, int c) {
  T v = (c < n_local) ? ld_x(x + c) : ld_x(xh + (c - n_local));
  if (CONJ_IN) v = conj_of(v);
  return v;
}

template <typename T, int EPI>
__device__ __forceinline__ void epilogue_acc(T acc, const T* w, int64_t r, T& e0, T& e1) {
  if (EPI == EPI_DOT_WY) {
    e0 = add(e0, mul(conj_of(w[r]), acc));  // conj_dot(w, y): src/vecalg.rs:564-568
  } else if (EPI == EPI_TT_TR) {
    const T cy = conj_of(acc);
    e0 = add(e0, mul(cy, acc));   // conj_dot(t, t)
    e1 = add(e1, mul(cy, w[r]));  // conj_dot(t, r)
  }
}

template <typename T, typename IP, int THREADS, int TILE, int EPI, bool CONJ_IN>
__global__ void __launch_bounds__(THREADS)
spmv_tile_kernel(const SpmvArgs<T, IP> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* s_val = reinterpret_cast<T*>(smem_raw);
  int* s_col = reinterpret_cast<int*>(s_val + (TILE + 4));
  T* s_red = reinterpret_cast<T*>(s_col + (TILE + 4));  // 32 T

  constexpr int ITERS = TILE / (THREADS * 4) + 1;  // +1: up to 3 head elements before the tile
  const int tid = threadIdx.x;
  T e0 = zero_of<T>(), e1 = zero_of<T>();
  if (a.gate && *a.gate != a.gate_value) return;

  for (int64_t ti = blockIdx.x; ti < a.ntiles; ti += gridDim.x) {
    const int tile = a.tile_list ? a.tile_list[ti] : (int)ti;
    const int r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
    const IP s = a.indptr[r0], e = a.indptr[r1];
    const int64_t cnt = (int64_t)(e - s);
    if (cnt <= TILE) {
      // ---- stage cols / vals: aligned 128-bit streaming loads, all issued before any store
      const IP s4 = s & ~(IP)3;
      const int total = (int)(e - s4);
      const int* gc = a.cols + s4;
      const T* gv = a.vals + s4;
      int4 c4[ITERS];
      ValPack<T> v4[ITERS];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = (it * THREADS + tid) * 4;
        if (j < total) {
          c4[it] = ld_stream_int4(gc + j);
          v4[it].load(gv + j);
        }
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int j = (it * THREADS + tid) * 4;
        if (j < total) {
          *reinterpret_cast<int4*>(s_col + j) = c4[it];
          v4[it].store(s_val + j);
        }
      }
      __syncthreads();
      // ---- one thread per row, sequential fold in CSR order (src/mat.rs:100-105)
      for (int r = r0 + tid; r < r1; r += THREADS) {
        const int p0 = (int)(a.indptr[r] - s4), p1 = (int)(a.indptr[r + 1] - s4);
        T acc = zero_of<T>();
#pragma unroll 4
        for (int k = p0; k < p1; ++k) {
          const T xv = gather_x<T, CONJ_IN>(a.x, a.xh, a.n_local, s_col[k]);
          acc = add(acc, mul(xv, s_val[k]));
        }
        a.y[r] = acc;
        epilogue_acc<T, EPI>(acc, a.w, r, e0, e1);
      }
      __syncthreads();
    } else {
      // ---- the tile holds a row longer than the staging buffer (rare).  Rows up to TILE/2
      //      non-zeros keep the sequential thread-per-row fold, read straight from global
      //      memory; longer rows are strided over by the whole CTA (their summation order
      //      differs from the reference).
      for (int r = r0 + tid; r < r1; r += THREADS) {
        const IP p0 = a.indptr[r], p1 = a.indptr[r + 1];
        if (p1 - p0 > (IP)(TILE / 2)) continue;
        T acc = zero_of<T>();
        for (IP k = p0; k < p1; ++k)
          acc = add(acc, mul(gather_x<T, CONJ_IN>(a.x, a.xh, a.n_local, a.cols[k]), a.vals[k]));
        a.y[r] = acc;
        epilogue_acc<T, EPI>(acc, a.w, r, e0, e1);
      }
      for (int r = r0; r < r1; ++r) {
        const IP p0 = a.indptr[r], p1 = a.indptr[r + 1];
        if (p1 - p0 <= (IP)(TILE / 2)) continue;
        T acc = zero_of<T>();
        for (IP k = p0 + tid; k < p1; k += THREADS)
          acc = add(acc, mul(gather_x<T, CONJ_IN>(a.x, a.xh, a.n_local, a.cols[k]), a.vals[k]));
        acc = block_sum(acc, s_red);
        if (tid == 0) {
          a.y[r] = acc;
          epilogue_acc<T, EPI>(acc, a.w, r, e0, e1);
        }
      }
    }
  }
  if (EPI != EPI_NONE) {
    e0 = block_sum(e0, s_red);
    if (EPI == EPI_TT_TR) e1 = block_sum(e1, s_red);
    if (tid == 0) {
      a.partials[2 * blockIdx.x] = e0;
      a.partials[2 * blockIdx.x + 1] = e1;
    }
  }
}

// ---------------------------------------------------------------- analysis kernels
template <typename IP>
__global__ void tile_rows_kernel(const IP* indptr, int64_t n, int64_t span, int64_t ntiles,
                                 int* tile_row) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  if (t == ntiles) {
    tile_row[t] = (int)n;
    return;
  }
  // first row r in [0, n] with indptr[r] >= t*span
  const int64_t target = t * span;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)indptr[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  tile_row[t] = (int)lo;
}

template <typename IP>
__global__ void max_row_kernel(const IP* indptr, int64_t n, unsigned long long* out) {
  unsigned long long m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long l = (unsigned long long)(indptr[i + 1] - indptr[i]);
    m = l > m ? l : m;
  }
  for (int d = 16; d > 0; d >>= 1) {
    const unsigned long long o = __shfl_down_sync(0xffffffffu, m, d);
    m = o > m ? o : m;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);  // integer atomic: order-independent
}

template <typename T>
__global__ void finalize_partials_kernel(const T* partials, int64_t nblocks, scal2* red) {
  __shared__ T scratch[32];
  for (int slot = 0; slot < 2; ++slot) {
    const T s = block_sum_partials(partials + slot, nblocks, 2, scratch);
    if (threadIdx.x == 0) red[slot] = to_scal2(s);
  }
}

// ---------------------------------------------------------------- configuration table
struct SpmvCfg {
  int threads, tile;
};
static const SpmvCfg kCfgs[] = {{128, 2048}, {256, 4096}, {128, 1024}, {256, 2048}};
static const int kNumCfgs = 4;

template <typename T>
static size_t spmv_smem_bytes(int tile) {
  return (size_t)(tile + 4) * (sizeof(T) + 4) + 32 * sizeof(T);
}

template <typename T, typename IP, int THREADS, int TILE>
static void launch_cfg(Ctx* ctx, const SpmvArgs<T, IP>& args, int epi, bool conj_in, int grid) {
  const size_t smem = spmv_smem_bytes<T>(TILE);
#define SPB_SPMV_CASE(E, C)                                                                      \
  {                                                                                              \
    auto k = spmv_tile_kernel<T, IP, THREADS, TILE, E, C>;                                       \
    static bool attr_set = false;                                                                \
    if (!attr_set) {                                                                             \
      SPB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_set = true;                                                                           \
    }                                                                                            \
    k<<<grid, THREADS, smem, ctx->stream>>>(args);                                               \
  }
  if (!conj_in) {
    if (epi == EPI_NONE) SPB_SPMV_CASE(EPI_NONE, false)
    else if (epi == EPI_DOT_WY) SPB_SPMV_CASE(EPI_DOT_WY, false)
    else SPB_SPMV_CASE(EPI_TT_TR, false)
  } else {
    if (epi == EPI_NONE) SPB_SPMV_CASE(EPI_NONE, true)
    else if (epi == EPI_DOT_WY) SPB_SPMV_CASE(EPI_DOT_WY, true)
    else SPB_SPMV_CASE(EPI_TT_TR, true)
  }
#undef SPB_SPMV_CASE
  check_launch("spmv_tile_kernel");
}

template <typename T, typename IP>
static void launch_spmv(Ctx* ctx, int cfg, const SpmvArgs<T, IP>& args, int epi, bool conj_in,
                        int grid) {
  LaunchScope ls(ctx, FAM_SPMV);
  switch (cfg) {
    case 0: launch_cfg<T, IP, 128, 2048>(ctx, args, epi, conj_in, grid); break;
    case 1: launch_cfg<T, IP, 256, 4096>(ctx, args, epi, conj_in, grid); break;
    case 2: launch_cfg<T, IP, 128, 1024>(ctx, args, epi, conj_in, grid); break;
    default: launch_cfg<T, IP, 256, 2048>(ctx, args, epi, conj_in, grid); break;
  }
}

template <typename T, typename IP>
static int spmv_blocks_per_sm(int cfg) {
  int nb = 0;
  const size_t smem = spmv_smem_bytes<T>(kCfgs[cfg].tile);
#define SPB_OCC(TH, TL)                                                                        \
  {                                                                                            \
    auto k = spmv_tile_kernel<T, IP, TH, TL, EPI_TT_TR, false>;                                \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, TH, smem);                           \
  }
  switch (cfg) {
    case 0: SPB_OCC(128, 2048) break;
    case 1: SPB_OCC(256, 4096) break;
    case 2: SPB_OCC(128, 1024) break;
    default: SPB_OCC(256, 2048) break;
  }
#undef SPB_OCC
  return nb > 0 ? nb : 1;
}

// ---------------------------------------------------------------- CsrMat methods
template <typename T>
void CsrMat<T>::analyze() {
  Ctx* c = ctx;
  // longest row
  DevBuf mx;
  mx.alloc(sizeof(unsigned long long));
  SPB_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(unsigned long long), c->stream));
  if (n_local > 0) {
    LaunchScope ls(c, FAM_SCALAR);
    const int grid = (int)std::min<int64_t>(ceil_div(n_local, 256), 4096);
    if (ip64)
      max_row_kernel<int64_t><<<grid, 256, 0, c->stream>>>(bufptr<int64_t>(indptr), n_local,
                                                          bufptr<unsigned long long>(mx));
    else
      max_row_kernel<int32_t><<<grid, 256, 0, c->stream>>>(bufptr<int32_t>(indptr), n_local,
                                                          bufptr<unsigned long long>(mx));
    check_launch("max_row_kernel");
  }
  unsigned long long h = 0;
  SPB_CUDA(cudaMemcpyAsync(&h, mx.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  max_row = (int64_t)h;

  // configuration: short rows -> small CTAs / small tiles keep ~all threads busy in the row phase
  const double mean = n_local > 0 ? (double)nnz / (double)n_local : 0.0;
  cfg = mean <= 40.0 ? 0 : 1;
  if (const char* e = getenv("SPB_SPMV_CFG")) {
    const int v = atoi(e);
    if (v >= 0 && v < kNumCfgs) cfg = v;
  }
  const int tile = kCfgs[cfg].tile;
  span = (max_row <= tile / 2) ? (tile - max_row) : tile / 2;
  if (span < 1) span = 1;
  ntiles = nnz / span + 1;
  tile_row.alloc(sizeof(int) * (size_t)(ntiles + 1));
  {
    LaunchScope ls(c, FAM_SCALAR);
    const int grid = (int)ceil_div(ntiles + 1, 256);
    if (ip64)
      tile_rows_kernel<int64_t><<<grid, 256, 0, c->stream>>>(bufptr<int64_t>(indptr), n_local, span,
                                                            ntiles, bufptr<int>(tile_row));
    else
      tile_rows_kernel<int32_t><<<grid, 256, 0, c->stream>>>(bufptr<int32_t>(indptr), n_local, span,
                                                            ntiles, bufptr<int>(tile_row));
    check_launch("tile_rows_kernel");
  }
  const int bps = ip64 ? spmv_blocks_per_sm<T, int64_t>(cfg) : spmv_blocks_per_sm<T, int32_t>(cfg);
  const int64_t max_grid = (int64_t)c->sm_count * bps;
  partials.alloc(sizeof(T) * 2 * (size_t)(2 * max_grid + 2));
  red.alloc(sizeof(scal2) * 2);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->dist && n_halo > 0) classify_tiles(this);
}

template <typename T>
void CsrMat<T>::mul(const T* x, T* y, int epi_mode, const T* w, bool conj_in) {
  Ctx* c = ctx;
  const int bps = ip64 ? spmv_blocks_per_sm<T, int64_t>(cfg) : spmv_blocks_per_sm<T, int32_t>(cfg);
  const int64_t max_grid = (int64_t)c->sm_count * bps;
  auto run = [&](const int* list, int64_t nt, int64_t part_off) -> int64_t {
    if (nt <= 0) return 0;
    const int grid = (int)std::min<int64_t>(nt, max_grid);
    if (ip64) {
      SpmvArgs<T, int64_t> a{bufptr<int64_t>(indptr), bufptr<int>(cols), bufptr<T>(vals), bufptr<int>(tile_row),
                             list, nt, x, bufptr<T>(halo), (int)n_local, y, w,
                             bufptr<T>(partials) + 2 * part_off, c->gate, c->gate_value};
      launch_spmv<T, int64_t>(c, cfg, a, epi_mode, conj_in, grid);
    } else {
      SpmvArgs<T, int32_t> a{bufptr<int32_t>(indptr), bufptr<int>(cols), bufptr<T>(vals), bufptr<int>(tile_row),
                             list, nt, x, bufptr<T>(halo), (int)n_local, y, w,
                             bufptr<T>(partials) + 2 * part_off, c->gate, c->gate_value};
      launch_spmv<T, int32_t>(c, cfg, a, epi_mode, conj_in, grid);
    }
    return grid;
  };
  if (n_halo == 0 || !c->dist) {
    last_partial_blocks = run(nullptr, ntiles, 0);
  } else {
    // interior rows overlap the NVLink halo exchange; boundary rows wait for it
    halo_exchange_begin(this, x);
    const int64_t g0 = run(bufptr<int>(tiles_interior), n_tiles_interior, 0);
    halo_exchange_wait(this);
    const int64_t g1 = run(bufptr<int>(tiles_boundary), n_tiles_boundary, g0);
    last_partial_blocks = g0 + g1;
  }
}

template <typename T>
void CsrMat<T>::finalize_epilogue() {
  LaunchScope ls(ctx, FAM_SCALAR);
  finalize_partials_kernel<T><<<1, 256, 0, ctx->stream>>>(bufptr<T>(partials), last_partial_blocks,
                                                          bufptr<scal2>(red));
  check_launch("finalize_partials_kernel");
}

template struct CsrMat<double>;
template struct CsrMat<cplx>;

}  // namespace spb
