// gs_wave.cu -- block-wavefront Gauss-Seidel sweep (sweep body src/gauss_seidel.rs:111-125).
//
// A Gauss-Seidel sweep is a sparse triangular solve: x_i needs the x_j of the same sweep for the
// columns on one side of the diagonal (the PRODUCED side; the other triangle reads a vector that
// is complete before the sweep starts).  It is LATENCY-bound -- the time is the length of the
// longest dependency chain times the cost of one producer -> consumer hand-off -- so the design
// minimises the hand-off and the number of instructions between two hand-offs:
//
//   * rows are cut into blocks of `block_rows` consecutive rows, one CTA per block.  The x values
//     of the block live in SHARED MEMORY; a dependency inside the block is a shared-memory read
//     ordered by a named barrier between levels;
//   * every block walks its rows in GLOBAL level order (longest dependency chain over the whole
//     triangle), so all blocks advance along the same wavefront whatever the block boundaries;
//   * a dependency on another block is read from the output vector in global memory.  The vector
//     is pre-filled with a sentinel (a NaN payload no arithmetic produces) and every x_i is
//     published with ONE 8/16-byte store, so the value itself says that it is ready.  The polling
//     is done by two HELPER warps per CTA that run ahead of the row threads and drop each value
//     into a shared-memory slot of the ring stage (slots arrive sentinel-filled with the static
//     stream); a row thread therefore reads every x -- own block or not -- with one shared-memory
//     load at a byte offset that was resolved at analysis time, and only spins (on shared memory)
//     when the wavefront really has to wait for a neighbour;
//   * blocks are handed out by a ticket in sweep order, so a block only ever waits for blocks
//     that already run: no deadlock although the grid may exceed the number of resident CTAs;
//   * everything static is packed at analysis time in exactly the order the CTA consumes it --
//     per chunk an ELL slab (column-major, width padded to 4 with +0.0 * zero-slot entries, which
//     leave the sequential sum bit-identical because a sum that starts at +0.0 never is -0.0) --
//     and streamed through a shared-memory ring by one producer lane with 1-D bulk async copies
//     (TMA engine + mbarrier complete_tx): no DRAM latency on the dependency chain;
//   * the other triangle costs no latency either: a fully parallel pre-pass (which also permutes
//     rhs and sentinel-fills out) folds it per row when it precedes the produced entries in CSR
//     order (backward sweep: sigma starts from that prefix), or stores its products when it
//     follows them (forward sweep of the stationary solver: the row thread only adds them);
//   * one thread per row folds sigma sequentially in CSR order (src/gauss_seidel.rs:113-118) with
//     separate multiply and add (-fmad=false): every x_i is bit-identical to the reference loop.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ops.cuh"
#include "tma.cuh"
#include "vecops.cuh"

namespace spb {

static const int WAVE_NC = 256;                          // row threads
static const int WAVE_NH = 128;                          // helper threads (cross-block values)
static const int WAVE_HB = 4;                            // polls a helper thread keeps in flight
static const int WAVE_THREADS = WAVE_NC + WAVE_NH + 32;  // + one producer warp
static const int WAVE_MAX_STAGES = 8;
static const int WAVE_FIXED = 256;     // barriers, ticket slot, zero slot
static const int WAVE_ZERO_OFF = 192;  // 16 bytes of +0.0: target of the ELL padding entries
#define SPB_GS_SENTINEL 0xFFFFDEADBEEF5EEDULL

__host__ __device__ inline int wave_a16(long long v) { return (int)((v + 15) & ~15LL); }

// Byte offsets of the sections of one packed chunk (all 16-byte aligned).
// header (32 B): nrows, nseg, W, nhalo, Wo, 0, 0, 0
struct WaveLayout {
  int seg_end, rowid, diag, eoff, eval, hslot, hcol, total;
};
template <typename T>
__host__ __device__ inline WaveLayout wave_layout(int nrows, int nseg, int W, int nhalo) {
  WaveLayout L;
  int off = 32;
  L.seg_end = off;
  off += wave_a16(4LL * nseg);
  L.rowid = off;
  off += wave_a16(4LL * nrows);
  L.diag = off;
  off += wave_a16((long long)sizeof(T) * nrows);
  L.eoff = off;  // ELL, column-major: entry e of row q at [e * nrows + q]; smem byte offset of its x
  off += wave_a16(4LL * W * nrows);
  L.eval = off;
  off += wave_a16((long long)sizeof(T) * W * nrows);
  L.hslot = off;  // sentinel-filled landing slots of the cross-block values
  off += wave_a16((long long)sizeof(T) * nhalo);
  L.hcol = off;   // their global column ids
  off += wave_a16(4LL * nhalo);
  L.total = off;
  return L;
}

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ double wv_poll(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ cplx wv_poll(const cplx* p) {
  cplx v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.re), "=d"(v.im) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wv_is_sentinel(double v) {
  return (unsigned long long)__double_as_longlong(v) == SPB_GS_SENTINEL;
}
__device__ __forceinline__ bool wv_is_sentinel(cplx v) { return wv_is_sentinel(v.re) || wv_is_sentinel(v.im); }
__device__ __forceinline__ void wv_publish(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void wv_publish(cplx* p, cplx v) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.re), "d"(v.im) : "memory");
}
template <typename T>
__device__ __forceinline__ T wv_lds(uint32_t addr);
template <>
__device__ __forceinline__ double wv_lds<double>(uint32_t addr) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
template <>
__device__ __forceinline__ cplx wv_lds<cplx>(uint32_t addr) {
  cplx v;
  asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(addr) : "memory");
  return v;
}
// x value at a resolved shared-memory address; spins while the slot still holds the sentinel
// (a cross-block value the helper warps have not delivered yet).
template <typename T>
__device__ __forceinline__ T wv_x(uint32_t addr, int* flag, long long& spins) {
  T x = wv_lds<T>(addr);
  if (wv_is_sentinel(x)) {
    long long n = 0;
    do {
      x = wv_lds<T>(addr);
    } while (wv_is_sentinel(x) && ++n < (1LL << 26));
    if (wv_is_sentinel(x)) *flag = 1;  // a legitimate value that equals the sentinel (a NaN): use it
    spins += n;
  }
  return x;
}

template <typename T>
struct WaveArgs {
  const unsigned char* stat;
  const WaveChunk* chunks;
  const int* blk_chunk;
  const T* rhsp;
  const T* aux;  // null: the other triangle is skipped (sweep from zero)
  T* out;
  int* ticket;  // [0] block ticket, [1] timeout flag
  int nblocks, block_rows;
  int stages, stage_static, stage_rhs_bytes, stage_bytes;
  long long* stats;  // null, or [4 * nblocks]: clocks total / waiting for the ring / shared-memory spins / thread 0 in the level barrier
  const int* gate;
  int gate_value;
};

// Pre-pass (fully parallel): sentinel-fill out, permute rhs into sweep order, reduce the other
// triangle to what the sweep needs (see the header), reset the block ticket.
template <typename T, typename IP, bool BWD>
__global__ void __launch_bounds__(kVecThreads) gs_wave_prep_kernel(int64_t n8, unsigned long long* out8, int64_t rhs_slots, const int* rowmap,
                                                                    const T* rhs, T* rhsp, const T* other, const IP* indptr, const int* cols,
                                                                    const T* vals, T* aux, const long long* aux_base, const int* aux_dims,
                                                                    int* ticket, const int* gate, int gate_value) {
  if (gate && *gate != gate_value) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) ticket[0] = 0;
  SPB_GRID_STRIDE(i, n8) out8[i] = SPB_GS_SENTINEL;
  SPB_GRID_STRIDE(p, rhs_slots) {
    const int r = rowmap[p];
    rhsp[p] = r >= 0 ? rhs[r] : zero_of<T>();
    if (!other) continue;
    if (BWD) {  // lower entries come first in CSR order: fold them now (src/gauss_seidel.rs:113-118)
      T sigma = zero_of<T>();
      if (r >= 0)
        for (IP k = indptr[r]; k < indptr[r + 1]; ++k) {
          const int c = cols[k];
          if (c < r) sigma = add(sigma, mul(vals[k], other[c]));
        }
      aux[p] = sigma;
    } else if (r >= 0) {  // upper entries follow the produced ones: their products, ELL column-major
      const long long base = aux_base[p];
      const int stride = aux_dims[2 * p], wo = aux_dims[2 * p + 1];
      int e = 0;
      for (IP k = indptr[r]; k < indptr[r + 1]; ++k) {
        const int c = cols[k];
        if (c > r) {
          aux[base + (long long)e * stride] = mul(vals[k], other[c]);
          ++e;
        }
      }
      for (; e < wo; ++e) aux[base + (long long)e * stride] = zero_of<T>();
    }
  }
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(WAVE_THREADS, 1) gs_wave_kernel(const WaveArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (a.gate && *a.gate != a.gate_value) return;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + WAVE_MAX_STAGES;
  int* s_ticket = reinterpret_cast<int*>(smem + 2 * WAVE_MAX_STAGES * sizeof(uint64_t));
  T* xs = reinterpret_cast<T*>(smem + WAVE_FIXED);
  unsigned char* ring = smem + WAVE_FIXED + ((size_t)a.block_rows * sizeof(T) + 127) / 128 * 128;
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x;
  const long long t_start = a.stats ? clock64() : 0;
  if (tid == 0) {
    *s_ticket = atomicAdd(a.ticket, 1);
    *reinterpret_cast<T*>(smem + WAVE_ZERO_OFF) = zero_of<T>();
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (WAVE_NC + WAVE_NH) / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int t = *s_ticket;  // blocks are processed in ticket order: a block waits only for earlier tickets
  const int b = BWD ? a.nblocks - 1 - t : t;
  const int r0 = b * a.block_rows;
  const int c0 = a.blk_chunk[t], c1 = a.blk_chunk[t + 1];
  const int S = a.stages;

  if (tid >= WAVE_NC + WAVE_NH) {  // ---- producer warp: one lane streams the block's chunks through the ring
    if (tid == WAVE_NC + WAVE_NH) {
      const uint64_t pol = l2_policy_evict_first();
      WaveChunk dn = c0 < c1 ? a.chunks[c0] : WaveChunk{};
      for (int c = c0; c < c1; ++c) {
        const int k = c - c0, s = k % S, u = k / S;
        const WaveChunk d = dn;
        if (c + 1 < c1) dn = a.chunks[c + 1];  // descriptor of the next chunk: in flight during the wait
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        unsigned char* st = ring + (size_t)s * a.stage_bytes;
        const uint32_t rb = (uint32_t)wave_a16((long long)sizeof(T) * d.nrows);
        const uint32_t ab = a.aux ? (uint32_t)wave_a16((long long)sizeof(T) * d.aux_cnt) : 0u;
        mbar_arrive_expect_tx(&full[s], (uint32_t)d.sbytes + rb + ab);
        bulk_g2s(st, a.stat + d.soff, (uint32_t)d.sbytes, &full[s], pol);
        bulk_g2s(st + a.stage_static, a.rhsp + d.rhs_off, rb, &full[s], pol);
        if (ab) bulk_g2s(st + a.stage_static + a.stage_rhs_bytes, a.aux + d.aux_off, ab, &full[s], pol);
      }
    }
    return;
  }

  if (tid >= WAVE_NC) {  // ---- helper warps: deliver the cross-block values of each chunk into its slots
    const int htid = tid - WAVE_NC;
    for (int c = c0; c < c1; ++c) {
      const int k = c - c0, s = k % S, u = k / S;
      mbar_wait(&full[s], (uint32_t)(u & 1));
      unsigned char* st = ring + (size_t)s * a.stage_bytes;
      const int* hdr = reinterpret_cast<const int*>(st);
      const int nhalo = hdr[3];
      if (nhalo > 0) {
        const WaveLayout L = wave_layout<T>(hdr[0], hdr[1], hdr[2], nhalo);
        const int* hcol = reinterpret_cast<const int*>(st + L.hcol);
        T* hslot = reinterpret_cast<T*>(st + L.hslot);
        for (int h0 = htid; h0 < nhalo; h0 += WAVE_NH * WAVE_HB) {  // slots are sorted by need: round-robin keeps every thread early
          const T* src[WAVE_HB];
          unsigned pend = 0;
#pragma unroll
          for (int j = 0; j < WAVE_HB; ++j) {
            const int h = h0 + j * WAVE_NH;
            src[j] = a.out + (h < nhalo ? hcol[h] : 0);
          }
          T v[WAVE_HB];
#pragma unroll
          for (int j = 0; j < WAVE_HB; ++j)
            if (h0 + j * WAVE_NH < nhalo) v[j] = wv_poll(src[j]);
#pragma unroll
          for (int j = 0; j < WAVE_HB; ++j) {
            if (h0 + j * WAVE_NH < nhalo) {
              if (wv_is_sentinel(v[j]))
                pend |= 1u << j;
              else
                hslot[h0 + j * WAVE_NH] = v[j];
            }
          }
          long long n = 0;
          while (pend && ++n < (1LL << 22)) {  // producers of other blocks still on their way
#pragma unroll
            for (int j = 0; j < WAVE_HB; ++j) {
              if (pend & (1u << j)) {
                const T w = wv_poll(src[j]);
                if (!wv_is_sentinel(w)) {
                  hslot[h0 + j * WAVE_NH] = w;
                  pend &= ~(1u << j);
                }
              }
            }
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    }
    return;
  }

  // ---- row threads
  long long wait_clk = 0, spins = 0, bar_clk = 0;
  for (int c = c0; c < c1; ++c) {
    const int k = c - c0, s = k % S, u = k / S;
    {
      const long long w0 = a.stats ? clock64() : 0;
      mbar_wait(&full[s], (uint32_t)(u & 1));
      if (a.stats) wait_clk += clock64() - w0;
    }
    const unsigned char* st = ring + (size_t)s * a.stage_bytes;
    const int* hdr = reinterpret_cast<const int*>(st);
    const int nrows = hdr[0], nseg = hdr[1], W = hdr[2], Wo = hdr[4];
    const WaveLayout L = wave_layout<T>(nrows, nseg, W, hdr[3]);
    const int* seg_end = reinterpret_cast<const int*>(st + L.seg_end);
    const int* rowid = reinterpret_cast<const int*>(st + L.rowid);
    const T* diag = reinterpret_cast<const T*>(st + L.diag);
    const uint32_t* eoff = reinterpret_cast<const uint32_t*>(st + L.eoff);
    const T* eval = reinterpret_cast<const T*>(st + L.eval);
    const T* rhss = reinterpret_cast<const T*>(st + a.stage_static);
    const T* auxs = reinterpret_cast<const T*>(st + a.stage_static + a.stage_rhs_bytes);
    int sbeg = 0;
    int send = seg_end[0];
    for (int g = 0; g < nseg; ++g) {
      const int send_next = g + 1 < nseg ? seg_end[g + 1] : send;
      for (int i = sbeg + tid; i < send; i += WAVE_NC) {
        // CSR order (src/gauss_seidel.rs:113-118): lower entries, then upper entries
        T sigma = (BWD && a.aux) ? auxs[i] : zero_of<T>();
        const T rv = rhss[i], dv = diag[i];
        const int row = rowid[i];
        for (int e = 0; e < W; e += 4) {
          const uint32_t o0 = eoff[(e + 0) * nrows + i], o1 = eoff[(e + 1) * nrows + i];
          const uint32_t o2 = eoff[(e + 2) * nrows + i], o3 = eoff[(e + 3) * nrows + i];
          const T v0 = eval[(e + 0) * nrows + i], v1 = eval[(e + 1) * nrows + i];
          const T v2 = eval[(e + 2) * nrows + i], v3 = eval[(e + 3) * nrows + i];
          T x0 = wv_lds<T>(sbase + o0), x1 = wv_lds<T>(sbase + o1), x2 = wv_lds<T>(sbase + o2), x3 = wv_lds<T>(sbase + o3);
          if (wv_is_sentinel(x0) | wv_is_sentinel(x1) | wv_is_sentinel(x2) | wv_is_sentinel(x3)) {  // rare: wait for a neighbour block
            x0 = wv_x<T>(sbase + o0, a.ticket + 1, spins);
            x1 = wv_x<T>(sbase + o1, a.ticket + 1, spins);
            x2 = wv_x<T>(sbase + o2, a.ticket + 1, spins);
            x3 = wv_x<T>(sbase + o3, a.ticket + 1, spins);
          }
          sigma = add(sigma, mul(v0, x0));
          sigma = add(sigma, mul(v1, x1));
          sigma = add(sigma, mul(v2, x2));
          sigma = add(sigma, mul(v3, x3));
        }
        if (!BWD && a.aux)
          for (int e = 0; e < Wo; ++e) sigma = add(sigma, auxs[e * nrows + i]);
        const T x = divi(sub(rv, sigma), dv);  // src/gauss_seidel.rs:123
        xs[row - r0] = x;
        wv_publish(a.out + row, x);
      }
      const long long w0 = a.stats ? clock64() : 0;
      consumer_bar_sync(WAVE_NC);  // the level is complete: its x values are visible in xs
      if (a.stats) bar_clk += clock64() - w0;
      sbeg = send;
      send = send_next;
    }
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  if (a.stats) {
    if (spins) atomicAdd(reinterpret_cast<unsigned long long*>(a.stats + 4 * t + 2), (unsigned long long)spins);
    if (tid == 0) {
      a.stats[4 * t + 0] = clock64() - t_start;
      a.stats[4 * t + 1] = wait_clk;
      a.stats[4 * t + 3] = bar_clk;
    }
  }
}

// ---- analysis ----------------------------------------------------------------------------------
template <typename U>
static void put_bytes(std::vector<unsigned char>& v, size_t off, const U* src, size_t count) {
  if (count) memcpy(v.data() + off, src, sizeof(U) * count);
}

template <typename T>
void wave_build(CsrMat<T>* A, const std::vector<int64_t>& ip, const std::vector<int>& cols, const std::vector<T>& vals,
                bool backward, WaveSched& ws) {
  Ctx* c = A->ctx;
  const int64_t n = A->n_local;
  ws.ok = false;
  ws.backward = backward;
  if (n <= 0 || n >= ((int64_t)1 << 31) - 1 || getenv("SPB_GS_LEGACY")) return;
  auto env = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
  };
  ws.stages = std::max(2, std::min(WAVE_MAX_STAGES, env("SPB_GS_STAGES", 3)));
  ws.stage_static = wave_a16(std::max(1024, env("SPB_GS_STAGE_BYTES", 16384)));
  ws.stage_rows = std::max(8, env("SPB_GS_STAGE_ROWS", 256));
  ws.stage_other = std::max(ws.stage_rows, env("SPB_GS_STAGE_OTHER", 1024));
  const int rhs_bytes = wave_a16((long long)sizeof(T) * ws.stage_rows);
  const int oth_bytes = wave_a16((long long)sizeof(T) * ws.stage_other);
  const int stage_bytes = (ws.stage_static + rhs_bytes + oth_bytes + 127) / 128 * 128;
  int smem_max = 0;
  SPB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  const int64_t avail = (int64_t)smem_max - WAVE_FIXED - (int64_t)ws.stages * stage_bytes - 256;
  const int64_t rmax = avail / (int64_t)sizeof(T);
  if (rmax < 64) return;
  int64_t R = env("SPB_GS_BLOCK_ROWS", 0);
  if (R <= 0) R = std::max<int64_t>(256, ceil_div(n, c->sm_count));
  R = std::max<int64_t>(1, std::min(R, rmax));
  const int64_t nb = ceil_div(n, R);
  ws.block_rows = (int)R;
  ws.nblocks = (int)nb;
  const size_t ring_off = (size_t)WAVE_FIXED + ((size_t)R * sizeof(T) + 127) / 128 * 128;
  ws.smem_bytes = ring_off + (size_t)ws.stages * stage_bytes;

  const int64_t align_el = std::max<int64_t>(1, 16 / (int64_t)sizeof(T));  // elements per 16 bytes
  auto align_slots = [&](int64_t v) { return (v + align_el - 1) / align_el * align_el; };
  auto pad4 = [](int v) { return (v + 3) & ~3; };
  const unsigned long long sentinel = SPB_GS_SENTINEL;

  std::vector<int> lev(n, 0);
  std::vector<unsigned char> stat;
  std::vector<WaveChunk> chunks;
  std::vector<int> blk_chunk(nb + 1, 0), rowmap, aux_dims;
  std::vector<long long> aux_base;
  stat.reserve((size_t)(ip[n] * (sizeof(T) + 4) + n * (24 + sizeof(T))));
  rowmap.reserve((size_t)n + 2 * (size_t)nb);
  int64_t max_levels = 0, aux_slots = 0;

  // per-row counts: produced entries, other-side entries, produced entries outside the block
  std::vector<int> order, cnt;
  struct RowInfo {
    int np, no, nx;
  };
  auto row_info = [&](int64_t i, int64_t r0, int64_t r1, RowInfo& ri) {
    ri.np = ri.no = ri.nx = 0;
    bool seen_upper = false, okrow = true;
    for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
      const int j = cols[k];
      if (j == i) continue;
      if (j < 0 || j >= n) okrow = false;
      if (j < i) {
        if (seen_upper) okrow = false;  // a lower entry after an upper one: the CSR-order fold would change
      } else {
        seen_upper = true;
      }
      if (backward ? (j > i) : (j < i)) {
        ++ri.np;
        if (j < r0 || j >= r1) ++ri.nx;
      } else {
        ++ri.no;
      }
    }
    return okrow;
  };

  // GLOBAL levels (longest dependency chain over the whole triangle).
  {
    auto visit = [&](int64_t i) {
      int l = 0;
      for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
        const int j = cols[k];
        if (j >= 0 && j < n && (backward ? j > i : j < i)) l = std::max(l, lev[j] + 1);
      }
      lev[i] = l;
    };
    if (backward)
      for (int64_t i = n - 1; i >= 0; --i) visit(i);
    else
      for (int64_t i = 0; i < n; ++i) visit(i);
  }

  // open chunk
  std::vector<int> ch_rows, ch_segend;
  int ch_w = 0, ch_wo = 0, ch_nhalo = 0;
  bool seg_open = false;

  auto close_chunk = [&](int64_t r0, int64_t r1, int kblk) {
    if (ch_rows.empty()) return;
    if (seg_open) ch_segend.push_back((int)ch_rows.size());
    seg_open = false;
    const int nrows = (int)ch_rows.size(), nseg = (int)ch_segend.size();
    const int W = pad4(ch_w), Wo = backward ? 0 : ch_wo;
    const WaveLayout L = wave_layout<T>(nrows, nseg, W, ch_nhalo);
    const size_t base = stat.size();
    stat.resize(base + L.total, 0);
    const uint32_t stage_base = (uint32_t)(ring_off + (size_t)(kblk % ws.stages) * stage_bytes);
    WaveChunk d{};
    d.soff = (long long)base;
    d.sbytes = L.total;
    d.nrows = nrows;
    d.rhs_off = (long long)rowmap.size();
    if (backward) {
      d.aux_off = d.rhs_off;  // one pre-folded value per row, same slots as rhs
      d.aux_cnt = nrows;
    } else {
      d.aux_off = aux_slots;
      d.aux_cnt = Wo * nrows;
      aux_slots = align_slots(aux_slots + d.aux_cnt);
    }
    std::vector<uint32_t> eoff((size_t)W * nrows, (uint32_t)WAVE_ZERO_OFF);
    std::vector<T> ev((size_t)W * nrows, zero_of<T>()), dg(nrows), hs(ch_nhalo);
    std::vector<int> hcol(ch_nhalo);
    for (int h = 0; h < ch_nhalo; ++h) {
      unsigned long long w[2] = {sentinel, sentinel};
      memcpy(&hs[h], w, sizeof(T));
    }
    int nh = 0;
    for (int q = 0; q < nrows; ++q) {
      const int64_t i = ch_rows[q];
      int e = 0;
      T dv = zero_of<T>();
      for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
        const int j = cols[k];
        if (j == i) {
          dv = vals[k];  // the last diagonal entry wins, as in the loop of src/gauss_seidel.rs:119-121
          continue;
        }
        if (!(backward ? (j > i) : (j < i))) continue;  // other triangle: handled by the pre-pass
        uint32_t off;
        if (j >= r0 && j < r1) {
          off = (uint32_t)(WAVE_FIXED + (size_t)(j - r0) * sizeof(T));
        } else {
          hcol[nh] = j;
          off = stage_base + (uint32_t)L.hslot + (uint32_t)(nh * sizeof(T));
          ++nh;
        }
        eoff[(size_t)e * nrows + q] = off;
        ev[(size_t)e * nrows + q] = vals[k];
        ++e;
      }
      dg[q] = dv;
      rowmap.push_back((int)i);
      if (!backward) {
        aux_base.push_back(d.aux_off + q);
        aux_dims.push_back(nrows);
        aux_dims.push_back(Wo);
      }
    }
    while ((int64_t)rowmap.size() != align_slots((int64_t)rowmap.size())) {
      rowmap.push_back(-1);
      if (!backward) {
        aux_base.push_back(0);
        aux_dims.push_back(0);
        aux_dims.push_back(0);
      }
    }
    const int hdr[8] = {nrows, nseg, W, ch_nhalo, Wo, 0, 0, 0};
    put_bytes(stat, base, hdr, 8);
    put_bytes(stat, base + L.seg_end, ch_segend.data(), nseg);
    put_bytes(stat, base + L.rowid, ch_rows.data(), nrows);
    put_bytes(stat, base + L.diag, dg.data(), nrows);
    put_bytes(stat, base + L.eoff, eoff.data(), eoff.size());
    put_bytes(stat, base + L.eval, ev.data(), ev.size());
    put_bytes(stat, base + L.hslot, hs.data(), hs.size());
    put_bytes(stat, base + L.hcol, hcol.data(), hcol.size());
    chunks.push_back(d);
    ch_rows.clear();
    ch_segend.clear();
    ch_w = ch_wo = ch_nhalo = 0;
  };

  for (int64_t t = 0; t < nb; ++t) {
    const int64_t b = backward ? nb - 1 - t : t;
    const int64_t r0 = b * R, r1 = std::min(n, r0 + R);
    blk_chunk[t] = (int)chunks.size();
    int minl = lev[r0], maxl = lev[r0];
    for (int64_t i = r0; i < r1; ++i) {
      minl = std::min(minl, lev[i]);
      maxl = std::max(maxl, lev[i]);
    }
    max_levels = std::max<int64_t>(max_levels, maxl - minl + 1);
    const int nl = maxl - minl + 1;
    cnt.assign(nl + 1, 0);
    for (int64_t i = r0; i < r1; ++i) cnt[lev[i] - minl + 1]++;
    for (int l = 0; l < nl; ++l) cnt[l + 1] += cnt[l];
    order.resize(r1 - r0);
    {
      std::vector<int> cur(cnt.begin(), cnt.end() - 1);
      for (int64_t i = r0; i < r1; ++i) order[cur[lev[i] - minl]++] = (int)i;
    }
    int kblk = 0;
    for (int l = 0; l < nl; ++l) {
      for (int q = cnt[l]; q < cnt[l + 1]; ++q) {
        const int64_t i = order[q];
        RowInfo ri;
        if (!row_info(i, r0, r1, ri)) return;
        if (wave_layout<T>(1, 1, pad4(ri.np), ri.nx).total > ws.stage_static || ri.no > ws.stage_other) return;  // row too long for a stage
        const int nr1 = (int)ch_rows.size() + 1;
        const int nseg1 = (int)ch_segend.size() + 1;
        const int w1 = std::max(ch_w, ri.np), wo1 = std::max(ch_wo, ri.no);
        if (!ch_rows.empty() &&
            (nr1 > ws.stage_rows || (!backward && (int64_t)wo1 * nr1 > ws.stage_other) ||
             wave_layout<T>(nr1, nseg1, pad4(w1), ch_nhalo + ri.nx).total > ws.stage_static)) {
          close_chunk(r0, r1, kblk++);
        }
        ch_rows.push_back((int)i);
        ch_w = std::max(ch_w, ri.np);
        ch_wo = std::max(ch_wo, ri.no);
        ch_nhalo += ri.nx;
        seg_open = true;
      }
      if (seg_open) {  // end of a level: barrier point
        ch_segend.push_back((int)ch_rows.size());
        seg_open = false;
      }
    }
    close_chunk(r0, r1, kblk++);
    if (chunks.size() >= (size_t)1 << 31) return;
  }
  blk_chunk[nb] = (int)chunks.size();
  ws.nchunks = (int64_t)chunks.size();
  ws.rhs_slots = (int64_t)rowmap.size();
  ws.aux_slots = backward ? ws.rhs_slots : aux_slots;
  ws.local_levels_max = max_levels;

  ws.stat.alloc(stat.size() + 64);
  ws.chunks.alloc(sizeof(WaveChunk) * std::max<size_t>(chunks.size(), 1));
  ws.blk_chunk.alloc(sizeof(int) * (nb + 1));
  ws.rowmap.alloc(sizeof(int) * std::max<size_t>(rowmap.size(), 1));
  ws.rhsp.alloc(sizeof(T) * (rowmap.size() + 4));
  ws.aux.alloc(sizeof(T) * ((size_t)ws.aux_slots + 4));
  ws.ticket.alloc(sizeof(int) * 4);
  SPB_CUDA(cudaMemcpyAsync(ws.stat.p, stat.data(), stat.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.chunks.p, chunks.data(), sizeof(WaveChunk) * chunks.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.blk_chunk.p, blk_chunk.data(), sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.rowmap.p, rowmap.data(), sizeof(int) * rowmap.size(), cudaMemcpyHostToDevice, c->stream));
  if (!backward) {
    ws.aux_base.alloc(sizeof(long long) * std::max<size_t>(aux_base.size(), 1));
    ws.aux_dims.alloc(sizeof(int) * std::max<size_t>(aux_dims.size(), 1));
    SPB_CUDA(cudaMemcpyAsync(ws.aux_base.p, aux_base.data(), sizeof(long long) * aux_base.size(), cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(ws.aux_dims.p, aux_dims.data(), sizeof(int) * aux_dims.size(), cudaMemcpyHostToDevice, c->stream));
  }
  SPB_CUDA(cudaMemsetAsync(ws.ticket.p, 0, sizeof(int) * 4, c->stream));
  SPB_CUDA(cudaMemsetAsync(ws.rhsp.p, 0, ws.rhsp.bytes, c->stream));
  SPB_CUDA(cudaMemsetAsync(ws.aux.p, 0, ws.aux.bytes, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
  ws.ok = true;
}

template <typename T, typename IP>
static void wave_prep_launch(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, T* out) {
  Ctx* c = M->ctx;
  CsrMat<T>* A = M->A;
  const int64_t n = A->n_local;
  const int64_t n8 = n * (int64_t)(sizeof(T) / 8);
  const int64_t work = std::max(n8, ws.rhs_slots);
  LaunchScope lsc(c, FAM_PRECOND);
  auto kern = ws.backward ? gs_wave_prep_kernel<T, IP, true> : gs_wave_prep_kernel<T, IP, false>;
  kern<<<vec_grid(c, work), kVecThreads, 0, c->stream>>>(n8, reinterpret_cast<unsigned long long*>(out), ws.rhs_slots, bufptr<int>(ws.rowmap),
                                                          rhs, bufptr<T>(ws.rhsp), other, bufptr<IP>(A->indptr), bufptr<int>(A->cols),
                                                          bufptr<T>(A->vals), bufptr<T>(ws.aux), bufptr<long long>(ws.aux_base),
                                                          bufptr<int>(ws.aux_dims), bufptr<int>(ws.ticket), c->gate, c->gate_value);
  check_launch("gs_wave_prep_kernel");
}

template <typename T>
void wave_sweep(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, T* out) {
  Ctx* c = M->ctx;
  const int64_t n = M->A->n_local;
  if (n == 0) return;
  const int rhs_bytes = wave_a16((long long)sizeof(T) * ws.stage_rows);
  const int oth_bytes = wave_a16((long long)sizeof(T) * ws.stage_other);
  const int stage_bytes = (ws.stage_static + rhs_bytes + oth_bytes + 127) / 128 * 128;
  if (M->A->ip64)
    wave_prep_launch<T, int64_t>(M, ws, rhs, other, out);
  else
    wave_prep_launch<T, int32_t>(M, ws, rhs, other, out);
  WaveArgs<T> a{};
  a.stat = bufptr<unsigned char>(ws.stat);
  a.chunks = bufptr<WaveChunk>(ws.chunks);
  a.blk_chunk = bufptr<int>(ws.blk_chunk);
  a.rhsp = bufptr<T>(ws.rhsp);
  a.aux = other ? bufptr<T>(ws.aux) : nullptr;
  a.out = out;
  a.ticket = bufptr<int>(ws.ticket);
  a.nblocks = ws.nblocks;
  a.block_rows = ws.block_rows;
  a.stages = ws.stages;
  a.stage_static = ws.stage_static;
  a.stage_rhs_bytes = rhs_bytes;
  a.stage_bytes = stage_bytes;
  a.stats = M->wave_stats.p ? bufptr<long long>(M->wave_stats) : nullptr;
  a.gate = c->gate;
  a.gate_value = c->gate_value;
  auto kern = ws.backward ? gs_wave_kernel<T, true> : gs_wave_kernel<T, false>;
  static size_t attr_set[2] = {0, 0};
  if (attr_set[ws.backward ? 1 : 0] < ws.smem_bytes) {
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.smem_bytes));
    attr_set[ws.backward ? 1 : 0] = ws.smem_bytes;
  }
  LaunchScope lsc(c, FAM_PRECOND);
  kern<<<ws.nblocks, WAVE_THREADS, ws.smem_bytes, c->stream>>>(a);
  check_launch("gs_wave_kernel");
}

#define SPB_INST_WAVE(T)                                                                                              \
  template void wave_build<T>(CsrMat<T>*, const std::vector<int64_t>&, const std::vector<int>&, const std::vector<T>&, \
                              bool, WaveSched&);                                                                      \
  template void wave_sweep<T>(GsOp<T>*, WaveSched&, const T*, const T*, T*);
SPB_INST_WAVE(double)
SPB_INST_WAVE(cplx)

}  // namespace spb
