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
//   * per tile, one producer lane streams col_idx / values / the indptr slice into a ring of
//     shared-memory stages with 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx: the
//     TMA engine, UBLKCP in SASS) -- the matrix is read exactly once, never touches L1 or
//     registers, and the next tiles are in flight while the current one is being reduced;
//   * the consumer warps run one thread per row accumulating  acc = acc + x[col] * val
//     SEQUENTIALLY IN CSR ORDER -- the same left fold as src/mat.rs:100-105 -- so y is
//     bit-identical to the reference; lanes of a warp walk the same stencil diagonal, so the x
//     gathers (ld.global.nc through L1, L2 evict_last) coalesce, and are issued 8 per batch;
//   * dot-product epilogues (<r0,v>, <t,t>, <t,r>, conj(x).y) are folded in, so the Krylov loops
//     never re-read the SpMV output for a reduction.
// Algorithmic bytes per launch: nnz*(sizeof(T)+4) + (n+1)*sizeof(indptr) + 2*n*sizeof(T).
#include <cub/cub.cuh>

#include <algorithm>
#include <type_traits>
#include <cmath>
#include <utility>
#include <vector>

#include "csr.cuh"
#include "dist.cuh"
#include "reduce.cuh"

#include "tma.cuh"
#include "vecops.cuh"

namespace spb {

// ---------------------------------------------------------------- x gathers (read-only path, L1)
// L2 residency control: the matrix is streamed exactly once (evict_first), every x entry is
// gathered nnz-per-column times over a window of a few grid planes (evict_last).  Without the
// hints the 12 bytes/nnz matrix stream pushes x out of the 126 MB L2 before the next plane reuses
// it (27-point 512^3: one plane of rows streams 91 MB of matrix) and the gathers pay HBM latency.
__device__ __forceinline__ double ld_x(const double* p, uint64_t pol) {
  double v;
  asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ cplx ld_x(const cplx* p, uint64_t pol) {
  cplx v;
  asm("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.re), "=d"(v.im) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_x(const float* p, uint64_t pol) {
  float v;
  asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ cplxf ld_x(const cplxf* p, uint64_t pol) {
  cplxf v;
  asm("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.re), "=f"(v.im) : "l"(p), "l"(pol));
  return v;
}

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
  Acc<T>* partials;  // [2 * gridDim.x] double-double partial sums of the epilogue
  const int* gate;  // optional solver gate (see Ctx::gate)
  int gate_value;
  int tile;      // staging capacity in non-zeros (multiple of 4)
  int rcap;      // staging capacity in indptr entries
  int stages;
  // peer transport (fused interior + boundary launch): the halo payload is double buffered by the
  // parity of hhead->seq, tiles [first_boundary, ntiles) of the list need the neighbours' puts
  HaloHead* hhead;
  long long halo_stride;
  long long first_boundary;
  // column-offset dictionary (DICT kernels): row r has columns r + doff[pid[r] * dict_w + k]
  const unsigned* rinfo;  // DICT: per row (pattern id) | (low 16 bits of indptr[row]) << 16, n + 1 entries
  const int* doff;
  int dict_w;
  int* err;  // Ctx::dev_err: set when a neighbour's halo flag never arrived
  // x window (DICT kernels, xw_nseg > 0): per tile the x entries of every run of consecutive column
  // offsets are one contiguous segment; the producer stages the segments in shared memory with bulk
  // copies and row r reads entry k at window element soff[pid[r] * dict_w + k] + (r - r0)
  const int* soff;
  int xw_nseg, xw_rows, xw_elems;
  int xw_gmin[kMaxXwinSegs], xw_pad[kMaxXwinSegs], xw_extra[kMaxXwinSegs], xw_start[kMaxXwinSegs];
};

// x entry for local column id c.  Single GPU (HALO = false): every column is owned.  Partitioned
// (HALO = true): ids >= n_local live in the halo buffer; xh_adj = xh - n_local, so both cases are
// base + c with a selected base -- one unconditional load (a branch around the load would make
// ptxas wait for each gather before issuing the next).
template <typename T, bool CONJ_IN, bool HALO>
__device__ __forceinline__ T gather_x(const T* x, const T* xh_adj, int n_local, int c, uint64_t pol) {
  const T* base = x;
  if (HALO) base = (c < n_local) ? x : xh_adj;
  T v = ld_x(base + c, pol);
  if (CONJ_IN) v = conj_of(v);
  return v;
}

template <typename T, int EPI>
__device__ __forceinline__ void epilogue_acc(T acc, const T* w, int64_t r, Acc<T>& e0, Acc<T>& e1) {
  if (EPI == EPI_DOT_WY) {
    acc_prod(e0, conj_of(w[r]), acc);  // conj_dot(w, y): src/vecalg.rs:564-568
  } else if (EPI == EPI_TT_TR) {
    const T cy = conj_of(acc);
    acc_prod(e0, cy, acc);   // conj_dot(t, t)
    acc_prod(e1, cy, w[r]);  // conj_dot(t, r)
  }
}
// Same with the epilogue operand w[r] already in a register: the streaming row loops load it BEFORE the
// fold, so its global-memory latency hides behind the row instead of delaying the release of the stage.
template <typename T, int EPI>
__device__ __forceinline__ void epilogue_acc_v(T acc, T wv, Acc<T>& e0, Acc<T>& e1) {
  if (EPI == EPI_DOT_WY) {
    acc_prod(e0, conj_of(wv), acc);
  } else if (EPI == EPI_TT_TR) {
    const T cy = conj_of(acc);
    acc_prod(e0, cy, acc);
    acc_prod(e1, cy, wv);
  }
}

struct TileMeta {
  int r0, r1;     // rows of the tile
  int total;      // staged non-zeros counted from the 4-aligned base; -1 => long-row tile
  int ip_off;     // index of indptr[r0] inside the staged indptr slice; -1 => slice not staged
  long long s4;   // 4-aligned first non-zero
  int win;        // the stage carries the x window of this tile
  int pid_off;    // index of pid[r0] inside the staged pattern-id slice; -1 => not staged
  int halo;       // the tile may gather halo columns (a boundary tile of a partitioned matrix)
};

__host__ __device__ inline int align16i(int v) { return (v + 15) & ~15; }

// Sum over the consumer threads only (named barrier 1); result valid in consumer thread 0.
template <typename T>
__device__ __forceinline__ T consumer_sum(T v, T* scratch, int nthreads) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[wid] = v;
  consumer_bar_sync(nthreads);
  T tot = zero_of<T>();
  if (threadIdx.x == 0)
    for (int w = 0; w < nthreads / 32; ++w) tot = add(tot, scratch[w]);
  consumer_bar_sync(nthreads);
  return tot;
}

static const int kMaxStages = 4;

template <typename T, typename IP>
__device__ __forceinline__ void row_range(const SpmvArgs<T, IP>& a, const TileMeta& m, const IP* s_ip, int r,
                                          int& p0, int& p1) {
  if (m.ip_off >= 0) {
    const int q = m.ip_off + (r - m.r0);
    p0 = (int)((long long)s_ip[q] - m.s4);
    p1 = (int)((long long)s_ip[q + 1] - m.s4);
  } else {
    p0 = (int)((long long)a.indptr[r] - m.s4);
    p1 = (int)((long long)a.indptr[r + 1] - m.s4);
  }
}

// DICT kernels: extent and pattern of row r from the staged row words -- the low 16 bits of the row pointers are enough
// inside a tile (a stage holds < 65536 entries), so the 4- or 8-byte indptr stream is not read at all.
template <typename T, typename IP>
__device__ __forceinline__ void dict_row(const SpmvArgs<T, IP>& a, const TileMeta& m, const unsigned* s_info, int r, int& p0, int& p1,
                                         int& pidv) {
  if (m.pid_off >= 0) {
    const unsigned i0 = s_info[m.pid_off + (r - m.r0)], i1 = s_info[m.pid_off + (r - m.r0) + 1];
    const unsigned base = (unsigned)m.s4 & 0xFFFFu;
    p0 = (int)(((i0 >> 16) - base) & 0xFFFFu);
    p1 = (int)(((i1 >> 16) - base) & 0xFFFFu);
    pidv = (int)(i0 & 0xFFFFu);
  } else {  // a tile with more rows than the staging capacity: global memory
    p0 = (int)((long long)a.indptr[r] - m.s4);
    p1 = (int)((long long)a.indptr[r + 1] - m.s4);
    pidv = (int)(a.rinfo[r] & 0xFFFFu);
  }
}

// blockDim.x = CT consumer threads + one producer warp.
template <typename T, typename IP, bool HALO, int EPI, bool CONJ_IN, bool DICT>
__global__ void __launch_bounds__(288)
spmv_tma_kernel(const SpmvArgs<T, IP> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int CT = (int)blockDim.x - 32;
  const int VAL_BYTES = align16i((a.tile + 4) * (int)sizeof(T));
  const int COL_BYTES = DICT ? 0 : align16i((a.tile + 4) * 4);  // DICT: the column stream is not read at all
  const int IP_BYTES = DICT ? 0 : align16i((a.rcap + 8) * (int)sizeof(IP));  // DICT: row extents come with the pattern word
  const int XW_BYTES = DICT ? align16i(a.xw_elems * (int)sizeof(T)) : 0;
  const int PID_BYTES = DICT ? align16i((a.rcap + 16) * 4) : 0;  // the tile's row words: pattern id + low 16 bits of the row pointer
  const int STAGE_BYTES = VAL_BYTES + COL_BYTES + IP_BYTES + XW_BYTES + PID_BYTES;  // vals | cols | indptr slice | x window | pattern ids
  const int STAGES = a.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + STAGES * STAGE_BYTES);
  uint64_t* empty = full + kMaxStages;
  TileMeta* meta = reinterpret_cast<TileMeta*>(empty + kMaxStages);
  Acc<T>* s_acc = reinterpret_cast<Acc<T>*>((reinterpret_cast<uintptr_t>(meta + kMaxStages) + 15) & ~(uintptr_t)15);
  T* s_red = reinterpret_cast<T*>(s_acc);  // (scratch of the long-row path; never live at the same time)

  if (a.gate && *a.gate != a.gate_value) return;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t my_tiles = a.ntiles > (int64_t)blockIdx.x
                               ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x
                               : 0;
  Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
  const T* xh = a.xh;
  unsigned long long hseq = 0;
  if (HALO && a.hhead) {  // the put kernel queued before this launch has set seq for this exchange
    hseq = a.hhead->seq;
    xh += (long long)(hseq & 1) * a.halo_stride;
  }

  if (tid >= CT) {
    // ------------------------------------------------------------ producer (one lane)
    if (tid == CT) {
      constexpr int IPV = 16 / (int)sizeof(IP);  // indptr entries per 16 bytes
      const uint64_t pol_stream = l2_policy_evict_first();
      const uint64_t pol_keep = l2_policy_evict_last();
      constexpr int XA = 16 / (int)sizeof(T) > 0 ? 16 / (int)sizeof(T) : 1;  // elements per 16 bytes
      int s = 0;
      uint32_t ph = 0;
      bool halo_ready = !(HALO && a.hhead);
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t ti = blockIdx.x + i * gridDim.x;
        if (HALO && !halo_ready && ti >= a.first_boundary) {
          // boundary tiles gather x entries the neighbours store into this rank's halo window:
          // wait for their flags (they were sent before the neighbours' own SpMV, i.e. about one
          // interior pass ago).  Consumers only touch a tile after the full-barrier below.
          for (int j = 0; j < a.hhead->npeers; ++j)
            if (!spin_until_ge(&a.hhead->flags[a.hhead->peer_rank[j]], hseq)) {
              a.hhead->error = 1;
              if (a.err) *a.err = 1;
            }
          halo_ready = true;
        }
        const int tile = a.tile_list ? a.tile_list[ti] : (int)ti;
        const int r0 = a.tile_row[tile], r1 = a.tile_row[tile + 1];
        const IP st = a.indptr[r0], en = a.indptr[r1];
        const IP s4 = st & ~(IP)3;
        mbar_wait(&empty[s], ph ^ 1u);  // stage released by every consumer warp
        TileMeta m;
        m.r0 = r0;
        m.r1 = r1;
        m.s4 = (long long)s4;
        m.halo = (HALO && (!a.hhead || ti >= a.first_boundary)) ? 1 : 0;  // (NCCL transport: separate launches, every listed tile)
        unsigned char* stage = smem_raw + s * STAGE_BYTES;
        if ((int64_t)(en - st) <= a.tile) {
          m.total = (int)(en - s4);
          const int ra = r0 & ~(IPV - 1);  // 16-byte aligned start of the indptr slice
          const int nip = ((r1 + 1 - ra) + IPV - 1) & ~(IPV - 1);
          const bool ip_ok = !DICT && nip <= a.rcap + 8 - IPV;
          m.ip_off = ip_ok ? (r0 - ra) : -1;
          const uint32_t groups = (uint32_t)((m.total + 3) >> 2);
          // x window: every segment must lie inside the owned part of x (tiles at the ends of the row
          // range and tiles with halo columns gather from global memory instead)
          const int rows = r1 - r0;
          uint32_t xw_bytes = 0;
          bool win = DICT && a.xw_nseg > 0 && groups > 0 && rows <= a.xw_rows && (r0 & (XA - 1)) == 0;
          if (win) {
            for (int g = 0; g < a.xw_nseg; ++g) {
              const long long s0 = (long long)r0 + a.xw_gmin[g] - a.xw_pad[g];
              const int len = (a.xw_extra[g] + rows + XA - 1) & ~(XA - 1);
              if (s0 < 0 || s0 + len > (long long)a.n_local) win = false;
              xw_bytes += (uint32_t)len * (uint32_t)sizeof(T);
            }
          }
          m.win = win ? 1 : 0;
          // row words of the tile's rows r0 .. r1 (one more than rows: the end of the last row): 4 per 16 bytes
          const int pa = r0 & ~3;
          const int npid = ((r1 + 1 - pa) + 3) & ~3;
          const bool pid_ok = DICT && groups > 0 && npid <= a.rcap + 12;
          m.pid_off = pid_ok ? (r0 - pa) : -1;
          meta[s] = m;
          const uint32_t bytes = groups * ((DICT ? 0u : 16u) + 4u * (uint32_t)sizeof(T)) + (ip_ok ? (uint32_t)nip * (uint32_t)sizeof(IP) : 0u) +
                                 (win ? xw_bytes : 0u) + (pid_ok ? (uint32_t)npid * 4u : 0u);
          if (bytes) {
            mbar_arrive_expect_tx(&full[s], bytes);
            if (groups) {
              bulk_g2s(stage, a.vals + s4, groups * 4u * (uint32_t)sizeof(T), &full[s], pol_stream);
              if (!DICT) bulk_g2s(stage + VAL_BYTES, a.cols + s4, groups * 16u, &full[s], pol_stream);
            }
            if (ip_ok) bulk_g2s(stage + VAL_BYTES + COL_BYTES, a.indptr + ra, (uint32_t)nip * (uint32_t)sizeof(IP), &full[s], pol_stream);
            if (pid_ok) bulk_g2s(stage + VAL_BYTES + COL_BYTES + IP_BYTES + XW_BYTES, a.rinfo + pa, (uint32_t)npid * 4u, &full[s], pol_stream);
            if (win) {
              unsigned char* xw = stage + VAL_BYTES + COL_BYTES + IP_BYTES;
              for (int g = 0; g < a.xw_nseg; ++g) {
                const long long s0 = (long long)r0 + a.xw_gmin[g] - a.xw_pad[g];
                const int len = (a.xw_extra[g] + rows + XA - 1) & ~(XA - 1);
                bulk_g2s(xw + (size_t)a.xw_start[g] * sizeof(T), a.x + s0, (uint32_t)len * (uint32_t)sizeof(T), &full[s], pol_keep);
              }
            }
          } else {
            mbar_arrive(&full[s]);
          }
        } else {
          m.total = -1;  // long-row tile: consumers read it from global memory
          m.ip_off = -1;
          m.win = 0;
          m.pid_off = -1;
          meta[s] = m;
          mbar_arrive(&full[s]);
        }
        if (++s == STAGES) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ consumers
    const uint64_t pol_x = l2_policy_evict_last();
    int s = 0;
    uint32_t ph = 0;
    for (int64_t i = 0; i < my_tiles; ++i) {
      mbar_wait(&full[s], ph);
      const TileMeta m = meta[s];
      const unsigned char* stage = smem_raw + s * STAGE_BYTES;
      const T* s_val = reinterpret_cast<const T*>(stage);
      const int* s_col = reinterpret_cast<const int*>(stage + VAL_BYTES);
      const IP* s_ip = reinterpret_cast<const IP*>(stage + VAL_BYTES + COL_BYTES);
      const unsigned* s_info = reinterpret_cast<const unsigned*>(stage + VAL_BYTES + COL_BYTES + IP_BYTES + XW_BYTES);
      if (DICT && m.total >= 0 && m.win) {
        // x window: every operand of the tile is in shared memory -- values from the stream, x from the
        // staged segments; the same sequential fold in CSR order, no global load on the critical path
        const T* s_xw = reinterpret_cast<const T*>(stage + VAL_BYTES + COL_BYTES + IP_BYTES);
        for (int r = m.r0 + tid; r < m.r1; r += CT) {
          int p0, p1, pidv;
          dict_row(a, m, s_info, r, p0, p1, pidv);
          T acc = zero_of<T>();
          T wv = zero_of<T>();
          if (EPI != EPI_NONE) wv = a.w[r];
          const int lr = r - m.r0;
          const int4* dp = reinterpret_cast<const int4*>(a.soff + pidv * a.dict_w);
          int k = p0;
          for (; k + 8 <= p1; k += 8) {
            const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
            dp += 2;
            T xv[8];
            xv[0] = s_xw[lr + q0.x]; xv[1] = s_xw[lr + q0.y]; xv[2] = s_xw[lr + q0.z]; xv[3] = s_xw[lr + q0.w];
            xv[4] = s_xw[lr + q1.x]; xv[5] = s_xw[lr + q1.y]; xv[6] = s_xw[lr + q1.z]; xv[7] = s_xw[lr + q1.w];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = add(acc, mul(CONJ_IN ? conj_of(xv[j]) : xv[j], s_val[k + j]));
          }
          if (k < p1) {
            const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
            T xv[8];
            xv[0] = s_xw[lr + q0.x]; xv[1] = s_xw[lr + q0.y]; xv[2] = s_xw[lr + q0.z]; xv[3] = s_xw[lr + q0.w];
            xv[4] = s_xw[lr + q1.x]; xv[5] = s_xw[lr + q1.y]; xv[6] = s_xw[lr + q1.z]; xv[7] = s_xw[lr + q1.w];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (k + j < p1) acc = add(acc, mul(CONJ_IN ? conj_of(xv[j]) : xv[j], s_val[k + j]));
          }
          a.y[r] = acc;
          epilogue_acc_v<T, EPI>(acc, wv, e0, e1);
        }
      } else if (m.total >= 0) {
        // one thread per row: x gathers go to registers (ld.global.nc through L1, L2 evict_last),
        // 8 per batch, then the sequential fold in CSR order (src/mat.rs:100-105).  Full batches
        // carry no predicates; only the last (partial) batch of a row is clamped / predicated.
        // Partitioned matrices: only the BOUNDARY tiles have halo columns; the interior tiles (97 % of the rows of a
        // z-slab) run the loop without the owned / halo base select (3 extra instructions per gather).
        const T* xb = a.x;
        const T* xh_adj = HALO ? (xh - a.n_local) : a.x;
        const int nl = a.n_local;
        auto tile_rows = [&](auto halo_tile) {
        constexpr bool HT = decltype(halo_tile)::value;
        for (int r = m.r0 + tid; r < m.r1; r += CT) {
          int p0, p1, pidv = 0;
          if (DICT)
            dict_row(a, m, s_info, r, p0, p1, pidv);
          else
            row_range(a, m, s_ip, r, p0, p1);
          T acc = zero_of<T>();
          T wv = zero_of<T>();
          if (EPI != EPI_NONE) wv = a.w[r];  // early: hidden behind the fold
          int k = p0;
          // DICT: the row's column offsets come from the pattern dictionary (a few KB, L1 resident;
          // neighbouring rows share the pattern, so the loads of a warp are broadcasts)
          // (dict_w is a multiple of 8 and padded with zero offsets: two 16-byte loads per batch,
          // also for the partial batch at the end of a row -- a padding entry gathers x[r])
          const int4* dp = nullptr;
          if (DICT) dp = reinterpret_cast<const int4*>(a.doff + pidv * a.dict_w);
          for (; k + 8 <= p1; k += 8) {
            int c[8];
            T xv[8];
            if (DICT) {
              const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
              dp += 2;
              c[0] = r + q0.x; c[1] = r + q0.y; c[2] = r + q0.z; c[3] = r + q0.w;
              c[4] = r + q1.x; c[5] = r + q1.y; c[6] = r + q1.z; c[7] = r + q1.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) c[j] = s_col[k + j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = gather_x<T, CONJ_IN, HT>(xb, xh_adj, nl, c[j], pol_x);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = add(acc, mul(xv[j], s_val[k + j]));
          }
          if (k < p1) {
            int c[8];
            T xv[8];
            if (DICT) {
              const int4 q0 = __ldg(dp), q1 = __ldg(dp + 1);
              c[0] = r + q0.x; c[1] = r + q0.y; c[2] = r + q0.z; c[3] = r + q0.w;
              c[4] = r + q1.x; c[5] = r + q1.y; c[6] = r + q1.z; c[7] = r + q1.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) c[j] = s_col[min(k + j, p1 - 1)];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = gather_x<T, CONJ_IN, HT>(xb, xh_adj, nl, c[j], pol_x);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (k + j < p1) acc = add(acc, mul(xv[j], s_val[k + j]));
          }
          a.y[r] = acc;
          epilogue_acc_v<T, EPI>(acc, wv, e0, e1);
        }
        };
        if (HALO && m.halo)
          tile_rows(std::true_type{});
        else
          tile_rows(std::false_type{});
      } else {
        // rows up to tile/2 non-zeros keep the sequential fold (read from global memory);
        // longer rows are strided over by all consumers (their summation order differs).
        for (int r = m.r0 + tid; r < m.r1; r += CT) {
          const IP p0 = a.indptr[r], p1 = a.indptr[r + 1];
          if (p1 - p0 > (IP)(a.tile / 2)) continue;
          T acc = zero_of<T>();
          for (IP k = p0; k < p1; ++k)
            acc = add(acc, mul(gather_x<T, CONJ_IN, HALO>(a.x, HALO ? (xh - a.n_local) : a.x, a.n_local, a.cols[k], pol_x), a.vals[k]));
          a.y[r] = acc;
          epilogue_acc<T, EPI>(acc, a.w, r, e0, e1);
        }
        for (int r = m.r0; r < m.r1; ++r) {
          const IP p0 = a.indptr[r], p1 = a.indptr[r + 1];
          if (p1 - p0 <= (IP)(a.tile / 2)) continue;
          T acc = zero_of<T>();
          for (IP k = p0 + tid; k < p1; k += CT)
            acc = add(acc, mul(gather_x<T, CONJ_IN, HALO>(a.x, HALO ? (xh - a.n_local) : a.x, a.n_local, a.cols[k], pol_x), a.vals[k]));
          acc = consumer_sum(acc, s_red, CT);
          if (tid == 0) {
            a.y[r] = acc;
            epilogue_acc<T, EPI>(acc, a.w, r, e0, e1);
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s
      if (++s == STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  if (EPI != EPI_NONE) {
    e0 = block_sum(e0, s_acc);
    if (EPI == EPI_TT_TR) e1 = block_sum(e1, s_acc);
    if (tid == 0) {
      a.partials[2 * blockIdx.x] = e0;
      a.partials[2 * blockIdx.x + 1] = e1;
    }
  }
}

// ---------------------------------------------------------------- analysis kernels
// align: tile boundaries are rounded down to a multiple of `align` rows (x window: bulk copies of x start
// at 16-byte boundaries)
template <typename IP>
__global__ void tile_rows_kernel(const IP* indptr, int64_t n, int64_t span, int64_t ntiles,
                                 int* tile_row, int align) {
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
  tile_row[t] = (int)(lo & ~(int64_t)(align - 1));
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

// ---------------------------------------------------------------- column-offset dictionary
// The part of the mkl_sparse_optimize analogue that shrinks the matrix stream: rows whose column
// offsets (col - row, in CSR order) coincide share a PATTERN.  A stencil-like matrix has a handful
// of patterns (27-point 512^3: one per combination of touched domain faces), so per row a 16-bit
// pattern id replaces nnz_row * 4 bytes of column indices -- the SpMV then streams 8 instead of 12
// bytes per non-zero (f64) and is bit-identical (same entries, same order).  Everything runs on the
// device (the 512^3 matrix never exists on the host): hash every row, stable radix sort of
// (hash, row), run heads -> pattern ids, a representative row per pattern fills the dictionary,
// and a verification pass compares every row with its dictionary entry (a hash collision, or too
// many / too long patterns, simply leaves the dictionary off).
__device__ __forceinline__ unsigned long long mix64(unsigned long long h, unsigned long long v) {
  h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
  h *= 0xBF58476D1CE4E5B9ULL;
  return h ^ (h >> 29);
}
template <typename IP>
__global__ void dict_hash_kernel(const IP* indptr, const int* cols, int64_t n, unsigned long long* keys, int* rows) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const IP p0 = indptr[r], p1 = indptr[r + 1];
    unsigned long long h = mix64(0x1234567ULL, (unsigned long long)(p1 - p0));
    for (IP k = p0; k < p1; ++k) h = mix64(h, (unsigned long long)(unsigned)(cols[k] - (int)r));
    keys[r] = h;
    rows[r] = (int)r;
  }
}
__global__ void dict_heads_kernel(const unsigned long long* keys, int64_t n, int* head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// head / run: flags and their exclusive scan over the sorted order
template <typename IP>
__global__ void dict_fill_kernel(const IP* indptr, const int* cols, const int* rows_sorted, const int* head, const int* run, int64_t n,
                                 int w, int* doff, int* dlen, unsigned short* pid) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = rows_sorted[i];
    const int u = run[i] + head[i] - 1;  // index of the run this row belongs to
    pid[r] = (unsigned short)u;
    if (head[i]) {  // representative of its pattern
      const IP p0 = indptr[r], p1 = indptr[r + 1];
      dlen[u] = (int)(p1 - p0);
      for (IP k = p0; k < p1; ++k) doff[(int64_t)u * w + (k - p0)] = cols[k] - r;
    }
  }
}
template <typename IP>
__global__ void dict_verify_kernel(const IP* indptr, const int* cols, int64_t n, int w, const int* doff, const int* dlen,
                                   const unsigned short* pid, int* bad) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const IP p0 = indptr[r], p1 = indptr[r + 1];
    const int u = pid[r];
    bool ok = dlen[u] == (int)(p1 - p0);
    for (IP k = p0; ok && k < p1; ++k) ok = doff[(int64_t)u * w + (k - p0)] == cols[k] - (int)r;
    if (!ok) atomicExch(bad, 1);
  }
}

// row word of the DICT kernels: pattern id | (low 16 bits of indptr[row]) << 16; entry n carries the end of the last row
template <typename IP>
__global__ void dict_pack_kernel(const IP* indptr, const unsigned short* pid, int64_t n, unsigned* rinfo) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x)
    rinfo[r] = (r < n ? (unsigned)pid[r] : 0u) | (((unsigned)indptr[r] & 0xFFFFu) << 16);
}

template <typename T, typename IP>
static void build_dict_impl(CsrMat<T>* m) {
  Ctx* c = m->ctx;
  const int64_t n = m->n_local;
  m->dict_on = false;
  m->xwin_on = false;
  const char* e = getenv("SPB_SPMV_DICT");
  if ((e && *e == '0') || n <= 0 || n >= ((int64_t)1 << 31) - 1 || m->max_row > 64 || m->max_row < 1) return;
  // Measured in round 1 (profiles/r01_spmv_dict.txt): WITHOUT the x window the dictionary pays when the
  // column stream is a third of the bytes and rows are long (27-point f64: 1.24x); 7-point rows and
  // complex128 values (4 of 20 bytes) gained nothing -- those kernels were bound by the x gathers.
  // With the x window the gathers are shared-memory reads, so the decision is taken after the runs of
  // the dictionary are known (below): window possible => dictionary on for every scalar type.
  const bool pays_without_window = sizeof(T) <= 8 && (double)m->nnz >= 12.0 * (double)n;
  // Measured (profiles/r02_xwin_sweep_*.txt): on B200 the window is NOT faster by default -- the 27-point
  // kernel at 384^3 runs at 6.28 TB/s (96 % of the copy peak) with L1/L2 gathers and at 4.97 TB/s with the
  // window (fewer resident CTAs, +34 % L2 -> SM traffic for the segments); the 7-point dictionary + window
  // kernel ties the plain stream (0.252 vs 0.256 ms at 256^3).  So it is opt-in: SPB_SPMV_XWIN=1, or
  // mv_hint(), whose timed candidates include it.
  const bool window_allowed = m->xwin_want;
  if (!(e && *e == '1') && !pays_without_window && !window_allowed) return;
  const int w = ((int)m->max_row + 7) & ~7;  // 8 offsets = two 16-byte loads per gather batch
  DevBuf keys, keys2, rows, rows2, head, run, tmp, bad;
  keys.alloc(8 * (size_t)n);
  keys2.alloc(8 * (size_t)n);
  rows.alloc(4 * (size_t)n);
  rows2.alloc(4 * (size_t)n);
  head.alloc(4 * (size_t)n);
  run.alloc(4 * (size_t)n);
  bad.alloc(16);
  SPB_CUDA(cudaMemsetAsync(bad.p, 0, 16, c->stream));
  const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)c->sm_count * 16);
  const IP* ip = bufptr<IP>(m->indptr);
  const int* cols = bufptr<int>(m->cols);
  {
    LaunchScope ls(c, FAM_SCALAR);
    dict_hash_kernel<IP><<<grid, 256, 0, c->stream>>>(ip, cols, n, keys.as<unsigned long long>(), rows.as<int>());
    check_launch("dict_hash_kernel");
  }
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), rows.as<int>(), rows2.as<int>(),
                                  (int)n, 0, 64, c->stream);
  tmp.alloc(tb);
  {
    LaunchScope ls(c, FAM_SCALAR);
    SPB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.as<unsigned long long>(), keys2.as<unsigned long long>(), rows.as<int>(),
                                             rows2.as<int>(), (int)n, 0, 64, c->stream));
    dict_heads_kernel<<<grid, 256, 0, c->stream>>>(keys2.as<unsigned long long>(), n, head.as<int>());
    check_launch("dict_heads_kernel");
  }
  size_t tb2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb2, head.as<int>(), run.as<int>(), (int)n, c->stream);
  tmp.ensure(tb2);
  {
    LaunchScope ls(c, FAM_SCALAR);
    SPB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, head.as<int>(), run.as<int>(), (int)n, c->stream));
  }
  int last_run = 0, last_head = 0;
  SPB_CUDA(cudaMemcpyAsync(&last_run, run.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaMemcpyAsync(&last_head, head.as<int>() + (n - 1), 4, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t u = (int64_t)last_run + last_head;
  if (u > 65535 || u * w > (1 << 18)) return;  // not a stencil-like matrix: keep the plain column stream
  DevBuf doff, dlen, pid;
  doff.alloc(sizeof(int) * (size_t)(u * w + 64));
  dlen.alloc(sizeof(int) * (size_t)u);
  pid.alloc(sizeof(unsigned short) * (size_t)n + 16);
  SPB_CUDA(cudaMemsetAsync(doff.p, 0, doff.bytes, c->stream));
  {
    LaunchScope ls(c, FAM_SCALAR);
    dict_fill_kernel<IP><<<grid, 256, 0, c->stream>>>(ip, cols, rows2.as<int>(), head.as<int>(), run.as<int>(), n, w, doff.as<int>(),
                                                      dlen.as<int>(), pid.as<unsigned short>());
    dict_verify_kernel<IP><<<grid, 256, 0, c->stream>>>(ip, cols, n, w, doff.as<int>(), dlen.as<int>(), pid.as<unsigned short>(),
                                                        bad.as<int>());
    check_launch("dict_verify_kernel");
  }
  int isbad = 0;
  SPB_CUDA(cudaMemcpyAsync(&isbad, bad.p, 4, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (isbad) return;  // hash collision: two different rows in one run
  // host copy of the (small) dictionary: the runs of consecutive offsets define the x window
  m->dict_off_host.assign((size_t)(u * w), 0);
  m->dict_len_host.assign((size_t)u, 0);
  SPB_CUDA(cudaMemcpyAsync(m->dict_off_host.data(), doff.p, sizeof(int) * (size_t)(u * w), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaMemcpyAsync(m->dict_len_host.data(), dlen.p, sizeof(int) * (size_t)u, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  {
    std::vector<int> offs;
    for (int64_t q = 0; q < u; ++q)
      for (int k = 0; k < m->dict_len_host[q]; ++k) offs.push_back(m->dict_off_host[q * w + k]);
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    int runs = offs.empty() ? 0 : 1;
    for (size_t k = 1; k < offs.size(); ++k) runs += offs[k] != offs[k - 1] + 1;
    const bool window_possible = window_allowed && runs >= 1 && runs <= kMaxXwinSegs;
    if (!(e && *e == '1') && !pays_without_window && !window_possible) return;  // keep the plain column stream
    m->xwin_on = window_possible;  // (the layout is fixed by build_plan, which may still drop it)
  }
  DevBuf rinfo;
  rinfo.alloc(sizeof(unsigned) * ((size_t)n + 1 + 16));
  SPB_CUDA(cudaMemsetAsync(rinfo.p, 0, rinfo.bytes, c->stream));
  {
    LaunchScope ls(c, FAM_SCALAR);
    dict_pack_kernel<IP><<<grid, 256, 0, c->stream>>>(ip, pid.as<unsigned short>(), n, rinfo.as<unsigned>());
    check_launch("dict_pack_kernel");
  }
  SPB_CUDA(cudaStreamSynchronize(c->stream));  // (pid goes out of scope)
  m->dict_off = std::move(doff);
  m->pid = std::move(rinfo);
  m->dict_w = w;
  m->dict_u = u;
  m->dict_on = true;
}

// ---------------------------------------------------------------- launch plumbing
template <typename T, typename IP>
static size_t spmv_smem_bytes(const CsrMat<T>* m) {
  const size_t stage = (size_t)align16i((m->plan_tile + 4) * (int)sizeof(T)) + (m->dict_on ? 0 : align16i((m->plan_tile + 4) * 4)) +
                       (m->dict_on ? 0 : align16i((m->plan_rcap + 8) * (int)sizeof(IP)));
  const size_t xw = m->dict_on && m->xwin_on ? (size_t)align16i(m->xwin_elems * (int)sizeof(T)) : 0;
  const size_t pidb = m->dict_on ? (size_t)align16i((m->plan_rcap + 16) * 4) : 0;
  return m->plan_stages * (stage + xw + pidb) + kMaxStages * (16 + sizeof(TileMeta)) + 32 * sizeof(Acc<T>) + 32;
}

// Runs f(kernel_pointer) for the kernel instance selected by (halo, epi, conj).
template <typename T, typename IP, bool HALO, bool DICT, typename F>
static void with_kernel_ax(int epi, bool conj_in, F&& f) {
  constexpr bool CZ = ScalarTraits<T>::is_complex;
  if (CZ && conj_in) {
    if (epi == EPI_NONE) f(spmv_tma_kernel<T, IP, HALO, EPI_NONE, CZ, DICT>);
    else if (epi == EPI_DOT_WY) f(spmv_tma_kernel<T, IP, HALO, EPI_DOT_WY, CZ, DICT>);
    else f(spmv_tma_kernel<T, IP, HALO, EPI_TT_TR, CZ, DICT>);
  } else {
    if (epi == EPI_NONE) f(spmv_tma_kernel<T, IP, HALO, EPI_NONE, false, DICT>);
    else if (epi == EPI_DOT_WY) f(spmv_tma_kernel<T, IP, HALO, EPI_DOT_WY, false, DICT>);
    else f(spmv_tma_kernel<T, IP, HALO, EPI_TT_TR, false, DICT>);
  }
}
template <typename T, typename IP, typename F>
static void with_kernel(int halo, bool dict, int epi, bool conj_in, F&& f) {
  if (halo) {
    if (dict) with_kernel_ax<T, IP, true, true>(epi, conj_in, f);
    else with_kernel_ax<T, IP, true, false>(epi, conj_in, f);
  } else {
    if (dict) with_kernel_ax<T, IP, false, true>(epi, conj_in, f);
    else with_kernel_ax<T, IP, false, false>(epi, conj_in, f);
  }
}

template <typename T, typename IP>
static void launch_spmv(CsrMat<T>* m, const SpmvArgs<T, IP>& args, int epi, bool conj_in, int grid) {
  Ctx* ctx = m->ctx;
  LaunchScope ls(ctx, FAM_SPMV);
  const size_t smem = spmv_smem_bytes<T, IP>(m);
  with_kernel<T, IP>(m->n_halo > 0 ? 1 : 0, m->dict_on, epi, conj_in, [&](auto kernel) {
    SPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, m->plan_ct + 32, smem, ctx->stream>>>(args);
  });
  check_launch("spmv_tma_kernel");
}

template <typename T, typename IP>
static int spmv_blocks_per_sm(CsrMat<T>* m) {
  int nb = 0;
  const size_t smem = spmv_smem_bytes<T, IP>(m);
  with_kernel<T, IP>(m->n_halo > 0 ? 1 : 0, m->dict_on, EPI_TT_TR, false, [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, m->plan_ct + 32, smem);
  });
  return nb > 0 ? nb : 1;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
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

  // ---- launch plan: a static rule (same plan for the same matrix on every run, so the fused
  //      dot products are bit-reproducible run to run); env overrides are for tuning sweeps.
  // One consumer thread per row and tile.  Measured on B200 (profiles/r01_spmv_plan_sweep.txt):
  // ~21 KB per stage is the sweet spot; rows of <= 12 non-zeros want a 2-stage ring (the row
  // phase is short, the bulk-copy latency dominates), longer rows want 1 stage and twice the
  // resident CTAs (the row phase dominates).
  {
    const char* xe = getenv("SPB_SPMV_XWIN");
    xwin_want = xe && *xe == '1';
  }
  choose_format();
  int ct, stages;
  static_plan(ct, stages);
  build_plan(ct, stages);
  partials.alloc(sizeof(Acc<T>) * 2 * (size_t)(2 * (int64_t)c->sm_count * 32 + 2));
  red.alloc(sizeof(scal2) * 2);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->dist && !peers.empty()) classify_tiles(this);
}

// Operand format of the stream: plain CSR, column-offset dictionary, dictionary + x window (xwin_want).
template <typename T>
void CsrMat<T>::choose_format() {
  if (ip64)
    build_dict_impl<T, int64_t>(this);
  else
    build_dict_impl<T, int32_t>(this);
}

// The static launch-plan rule for the chosen format.
template <typename T>
void CsrMat<T>::static_plan(int& ct, int& stages) {
  const double mean = n_local > 0 ? (double)nnz / (double)n_local : 0.0;
  const int mean_c = (int)std::max(1.0, std::ceil(mean));
  const int bytes_per_nnz = (int)sizeof(T) + (dict_on ? 0 : 4);
  ct = env_int("SPB_SPMV_CT", 0);
  if (ct <= 0) ct = (int)((21504 / ((int64_t)mean_c * bytes_per_nnz)) / 32 * 32);
  stages = env_int("SPB_SPMV_STAGES", 0);
  if (stages <= 0) stages = mean <= 12.0 ? 2 : 1;  // (also the best choice for the dictionary kernel: profiles/r01_spmv_dict.txt)
}

// Fix (consumer threads, stages), derive the tile size, cut the rows into tiles.
template <typename T>
void CsrMat<T>::build_plan(int ct, int stages) {
  Ctx* c = ctx;
  const double mean = n_local > 0 ? (double)nnz / (double)n_local : 0.0;
  const int mean_c = (int)std::max(1.0, std::ceil(mean));
  plan_gb = 0;
  plan_stages = std::max(1, std::min(kMaxStages, stages));
  const int rpt = std::max(1, env_int("SPB_SPMV_RPT", 1));
  const int64_t row_extra = std::min<int64_t>(max_row, 4096);
  plan_ct = std::max(32, std::min(256, ct / 32 * 32));
  auto stage_bytes = [&](int64_t tile) { return (tile + 4) * (int64_t)(sizeof(T) + (dict_on ? 0 : 4)) + (2 * 256 + 16) * (int64_t)sizeof(int64_t); };
  int64_t tile = std::max<int64_t>(256, (((int64_t)plan_ct * rpt * mean_c + row_extra) + 3) & ~3LL);
  const int64_t hard_cap = 200 * 1024;  // one CTA must fit
  while (plan_stages > 1 && plan_stages * stage_bytes(tile) > hard_cap) --plan_stages;
  while (tile > 256 && plan_stages * stage_bytes(tile) > hard_cap) tile = (tile / 2 + 3) & ~3LL;
  const int max_tile = env_int("SPB_SPMV_MAXTILE", 0);
  if (max_tile > 0) tile = std::min<int64_t>(tile, std::max(256, (max_tile + 3) & ~3));
  plan_tile = (int)tile;
  // indptr / pattern-id entries staged per tile: a typical tile has ~ct * rpt rows; tiles of shorter
  // (boundary) rows that exceed the capacity read their row pointers from global memory instead.  Kept
  // tight: with the 27-point plan one KB per stage decides between 8 and 9 resident CTAs per SM.
  plan_rcap = std::max(64, plan_ct * rpt + plan_ct * rpt / 4 + 16);
  span = (max_row <= plan_tile / 2) ? (plan_tile - max_row) : plan_tile / 2;
  if (span < 1) span = 1;
  // ---- x window: runs of consecutive column offsets -> one contiguous x segment per run and tile
  int tile_align = 1;
  if (dict_on && xwin_on) {
    const int XA = std::max<int>(1, 16 / (int)sizeof(T));  // elements per 16 bytes
    const int64_t span_w = (int64_t)plan_tile - (int64_t)XA * max_row;  // tile boundaries rounded down to XA rows
    std::vector<int> offs;
    for (int64_t q = 0; q < dict_u; ++q)
      for (int k = 0; k < dict_len_host[q]; ++k) offs.push_back(dict_off_host[q * dict_w + k]);
    std::sort(offs.begin(), offs.end());
    offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
    std::vector<std::pair<int, int>> runs;  // (first, last) offset
    for (int d : offs) {
      if (!runs.empty() && d == runs.back().second + 1)
        runs.back().second = d;
      else
        runs.emplace_back(d, d);
    }
    // rows of a typical tile: span non-zeros of mean-length rows (+ slack); shorter boundary rows give
    // tiles with more rows, which simply fall back to global gathers
    const int rows_typ = (int)std::min<int64_t>(4096, (int64_t)std::ceil((double)std::max<int64_t>(span_w, 1) / std::max(1.0, mean)) + 2 * XA + 8);
    xwin_rows = (rows_typ + XA - 1) / XA * XA;
    int acc = 0;
    bool ok = span_w >= plan_tile / 2 && (int)runs.size() <= kMaxXwinSegs && !runs.empty();
    if (ok) {
      xwin_nseg = (int)runs.size();
      for (int g = 0; g < xwin_nseg; ++g) {
        const int gmin = runs[g].first, gmax = runs[g].second;
        const int pad = ((gmin % XA) + XA) % XA;
        xwin_gmin[g] = gmin;
        xwin_pad[g] = pad;
        xwin_extra[g] = pad + (gmax - gmin);
        xwin_start[g] = acc;
        acc += (xwin_extra[g] + xwin_rows + XA - 1) / XA * XA;
      }
      xwin_elems = acc;
      ok = (int64_t)acc * (int64_t)sizeof(T) <= 32 * 1024 && plan_stages * (stage_bytes(tile) + (int64_t)acc * (int64_t)sizeof(T)) <= hard_cap;
    }
    if (ok) {
      std::vector<int> soff((size_t)(dict_u * dict_w) + 64, 0);
      for (int64_t q = 0; q < dict_u; ++q)
        for (int k = 0; k < dict_len_host[q]; ++k) {
          const int d = dict_off_host[q * dict_w + k];
          int g = 0;
          while (!(d >= runs[g].first && d <= runs[g].second)) ++g;
          soff[q * dict_w + k] = xwin_start[g] + xwin_pad[g] + (d - xwin_gmin[g]);
        }
      dict_soff.alloc(sizeof(int) * soff.size());
      SPB_CUDA(cudaMemcpyAsync(dict_soff.p, soff.data(), sizeof(int) * soff.size(), cudaMemcpyHostToDevice, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      span = span_w;
      tile_align = XA;
    } else {
      xwin_on = false;
      xwin_nseg = 0;
      xwin_elems = 0;
    }
  }
  ntiles = nnz / span + 1;
  tile_row.alloc(sizeof(int) * (size_t)(ntiles + 1));
  {
    LaunchScope ls(c, FAM_SCALAR);
    const int grid = (int)ceil_div(ntiles + 1, 256);
    if (ip64)
      tile_rows_kernel<int64_t><<<grid, 256, 0, c->stream>>>(bufptr<int64_t>(indptr), n_local, span,
                                                            ntiles, bufptr<int>(tile_row), tile_align);
    else
      tile_rows_kernel<int32_t><<<grid, 256, 0, c->stream>>>(bufptr<int32_t>(indptr), n_local, span,
                                                            ntiles, bufptr<int>(tile_row), tile_align);
    check_launch("tile_rows_kernel");
  }
  plan_bps = ip64 ? spmv_blocks_per_sm<T, int64_t>(this) : spmv_blocks_per_sm<T, int32_t>(this);
  const int bps_cap = env_int("SPB_SPMV_BPS", 0);
  if (bps_cap > 0) plan_bps = std::min(plan_bps, bps_cap);
  plan_bps = std::min(plan_bps, 32);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
}

// The mkl_sparse_set_mv_hint + mkl_sparse_optimize analogue (src/mkl_mat.rs:81-148): time a few
// (consumer threads, stages) plans with real SpMV launches on this matrix and keep the fastest.
// The plan fixes the tile boundaries and therefore the summation order of the fused dot-product
// epilogues; after tuning the choice is stable for the lifetime of the matrix.
template <typename T>
void CsrMat<T>::autotune() {
  Ctx* c = ctx;
  if (n_local == 0 || nnz == 0) return;
  static const int cand[][2] = {{64, 1}, {128, 1}, {256, 1}, {64, 2}, {128, 2}, {192, 2}, {256, 2}};
  DevBuf xb, yb;
  xb.alloc(sizeof(T) * (size_t)n_local);
  yb.alloc(sizeof(T) * (size_t)n_local);
  SPB_CUDA(cudaMemsetAsync(xb.p, 0, xb.bytes, c->stream));
  cudaEvent_t e0, e1;
  SPB_CUDA(cudaEventCreate(&e0));
  SPB_CUDA(cudaEventCreate(&e1));
  // candidates: {current format, format with the x window} x {static plan, the (threads, stages) grid}.
  // Whatever wins, the results are the same bits (every row is the CSR-order fold; exact epilogue sums).
  int best_fmt = -1, best_ct = plan_ct, best_st = plan_stages;
  float best_ms = 0.f;
  const bool keep_want = xwin_want;
  const int ncand = (int)(sizeof(cand) / sizeof(cand[0]));
  for (int fmt = 0; fmt < 2; ++fmt) {
    xwin_want = fmt == 1;
    choose_format();
    if (fmt == 1 && !(dict_on && xwin_on)) break;  // no window for this matrix
    int sct, sst;
    static_plan(sct, sst);
    for (int i = 0; i <= ncand; ++i) {
      const int ct = i < ncand ? cand[i][0] : sct, st = i < ncand ? cand[i][1] : sst;
      build_plan(ct, st);
      if (fmt == 1 && !xwin_on) continue;  // the window did not fit this plan
      if (c->dist && !peers.empty()) classify_tiles(this);
      mul(bufptr<T>(xb), bufptr<T>(yb), EPI_NONE, nullptr, false);
      SPB_CUDA(cudaEventRecord(e0, c->stream));
      for (int rep = 0; rep < 3; ++rep) mul(bufptr<T>(xb), bufptr<T>(yb), EPI_NONE, nullptr, false);
      SPB_CUDA(cudaEventRecord(e1, c->stream));
      SPB_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (best_fmt < 0 || ms < best_ms) {
        best_fmt = fmt;
        best_ct = ct;
        best_st = st;
        best_ms = ms;
      }
      if (fmt == 1) xwin_on = true;  // (build_plan clears it when the window does not fit; next candidate retries)
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  xwin_want = best_fmt == 1 || (best_fmt < 0 && keep_want);
  choose_format();
  build_plan(best_ct, best_st);
  if (c->dist && !peers.empty()) classify_tiles(this);
}

template <typename T>
void CsrMat<T>::mul(const T* x, T* y, int epi_mode, const T* w, bool conj_in) {
  Ctx* c = ctx;
  const int64_t max_grid = (int64_t)c->sm_count * plan_bps;
  HaloHead* hhead = nullptr;
  long long hstride = 0, first_boundary = 0;
  const T* halo_ptr = bufptr<T>(halo);
  if (halo_win) {
    hhead = static_cast<HaloHead*>(halo_win->local);
    hstride = std::max<int64_t>(n_halo, 1);
    first_boundary = n_tiles_interior;
    halo_ptr = reinterpret_cast<const T*>(static_cast<char*>(halo_win->local) + kHaloHeadBytes);
  }
  // x window of the dictionary kernels: the bulk copies need a 16-byte aligned x (a caller's odd slice is
  // still multiplied correctly, through the global gathers); the window stays reserved in the stage
  const bool window = dict_on && xwin_on && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  auto set_window = [&](auto& a) {
    a.soff = bufptr<int>(dict_soff);
    a.xw_nseg = window ? xwin_nseg : 0;
    a.xw_rows = xwin_rows;
    a.xw_elems = dict_on && xwin_on ? xwin_elems : 0;
    for (int g = 0; g < kMaxXwinSegs; ++g) {
      a.xw_gmin[g] = xwin_gmin[g];
      a.xw_pad[g] = xwin_pad[g];
      a.xw_extra[g] = xwin_extra[g];
      a.xw_start[g] = xwin_start[g];
    }
  };
  auto run = [&](const int* list, int64_t nt, int64_t part_off) -> int64_t {
    if (nt <= 0) return 0;
    const int grid = (int)std::min<int64_t>(nt, max_grid);
    if (ip64) {
      SpmvArgs<T, int64_t> a{bufptr<int64_t>(indptr), bufptr<int>(cols), bufptr<T>(vals), bufptr<int>(tile_row),
                             list, nt, x, halo_ptr, (int)n_local, y, w,
                             bufptr<Acc<T>>(partials) + 2 * part_off, c->gate, c->gate_value,
                             plan_tile, plan_rcap, plan_stages, hhead, hstride, first_boundary,
                             bufptr<unsigned>(pid), bufptr<int>(dict_off), dict_w, c->dev_err};
      set_window(a);
      launch_spmv<T, int64_t>(this, a, epi_mode, conj_in, grid);
    } else {
      SpmvArgs<T, int32_t> a{bufptr<int32_t>(indptr), bufptr<int>(cols), bufptr<T>(vals), bufptr<int>(tile_row),
                             list, nt, x, halo_ptr, (int)n_local, y, w,
                             bufptr<Acc<T>>(partials) + 2 * part_off, c->gate, c->gate_value,
                             plan_tile, plan_rcap, plan_stages, hhead, hstride, first_boundary,
                             bufptr<unsigned>(pid), bufptr<int>(dict_off), dict_w, c->dev_err};
      set_window(a);
      launch_spmv<T, int32_t>(this, a, epi_mode, conj_in, grid);
    }
    return grid;
  };
  if (!c->dist || peers.empty()) {
    last_partial_blocks = run(nullptr, ntiles, 0);
  } else if (halo_win) {
    // peer transport: put the boundary entries into the neighbours' windows, then ONE launch --
    // interior tiles first, the kernel itself waits for the neighbours' flags before its first
    // boundary tile (the NVLink transfer hides behind the interior pass)
    halo_put(this, x);
    last_partial_blocks = run(bufptr<int>(tiles_all), n_tiles_interior + n_tiles_boundary, 0);
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
void CsrMat<T>::finalize_epilogue(bool allreduce) {
  finalize_reduce<T>(ctx, bufptr<Acc<T>>(partials), last_partial_blocks, bufptr<scal2>(red), allreduce);
}

template <typename T>
CsrMat<T>::~CsrMat() {
  halo_release(this);
}

template struct CsrMat<double>;
template struct CsrMat<cplx>;
template struct CsrMat<float>;
template struct CsrMat<cplxf>;

}  // namespace spb
