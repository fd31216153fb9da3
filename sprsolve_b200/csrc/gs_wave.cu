// gs_wave.cu -- block-wavefront Gauss-Seidel sweep (sweep body src/gauss_seidel.rs:111-125).
//
// A Gauss-Seidel sweep is a sparse triangular solve: x_i needs the x_j of the same sweep for the
// columns on one side of the diagonal.  It is LATENCY-bound -- the time is the length of the
// longest dependency chain times the cost of one producer -> consumer hand-off -- so the design
// minimises the cost of a hand-off instead of chasing bandwidth:
//
//   * rows are cut into blocks of `block_rows` consecutive rows, one CTA per block.  The x values
//     of the block live in SHARED MEMORY; a dependency inside the block is a shared-memory read
//     ordered by a named barrier between LOCAL levels (~0.1 us per level instead of a trip
//     through L2 plus a grid barrier);
//   * a dependency on another block is read from the output vector in global memory.  The vector
//     is pre-filled with a sentinel (a NaN payload no arithmetic produces) and every x_i is
//     published with ONE 8/16-byte store, so the value itself says that it is ready: the consumer
//     re-loads it (L2, relaxed.gpu) until it is not the sentinel.  No flags, no fences;
//   * blocks are handed out by a ticket in sweep order, so a block only ever waits for blocks
//     that already run: no deadlock although the grid may exceed the number of resident CTAs;
//   * everything static a row needs (row id, extents, diagonal, values, pre-classified column
//     indices) is packed at analysis time in exactly the order the CTA consumes it and streamed
//     through a shared-memory ring by one producer lane with 1-D bulk async copies (TMA engine +
//     mbarrier complete_tx), so no DRAM latency sits on the dependency chain.  The per-apply
//     data (rhs, and the x entries of the triangle that does not depend on this sweep) is brought
//     into the same order by a fully parallel pre-pass and rides the same ring;
//   * one thread per row folds sigma sequentially in CSR order (src/gauss_seidel.rs:113-118), so
//     every x_i is bit-identical to the reference's sequential loop.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ops.cuh"
#include "tma.cuh"
#include "vecops.cuh"

namespace spb {

static const int WAVE_NC = 256;                 // consumer threads (one row each per pass)
static const int WAVE_THREADS = WAVE_NC + 32;   // + one producer warp
static const int WAVE_MAX_STAGES = 8;
static const int WAVE_FIXED = 256;              // barriers + ticket slot
static const int WAVE_PF = 8;                   // produced-side entries gathered per batch
#define SPB_GS_SENTINEL 0xFFFFDEADBEEF5EEDULL
#define SPB_WAVE_INTRA 0x80000000u

__host__ __device__ inline int wave_a16(long long v) { return (int)((v + 15) & ~15LL); }

// Byte offsets of the sections of one packed chunk (all 16-byte aligned).
struct WaveLayout {
  int seg_end, rowid, pptr, nprod, optr, diag, eval, ecol, total;
};
template <typename T>
__host__ __device__ inline WaveLayout wave_layout(int nrows, int nseg, int nent) {
  WaveLayout L;
  int off = 16;  // header: nrows, nseg, nent, 0
  L.seg_end = off;
  off += wave_a16(4LL * nseg);
  L.rowid = off;
  off += wave_a16(4LL * nrows);
  L.pptr = off;
  off += wave_a16(4LL * (nrows + 1));
  L.nprod = off;
  off += wave_a16(4LL * nrows);
  L.optr = off;
  off += wave_a16(4LL * nrows);
  L.diag = off;
  off += wave_a16((long long)sizeof(T) * nrows);
  L.eval = off;
  off += wave_a16((long long)sizeof(T) * nent);
  L.ecol = off;
  off += wave_a16(4LL * nent);
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

// Publish variants (A/B knob SPB_GS_PUBLISH): 0 = relaxed store, 1 = atomic exchange (performed at
// L2 at once), 2 = relaxed store + gpu-scope fence.
__device__ __forceinline__ void wv_publish_mode(double* p, double v, int mode) {
  if (mode == 1) {
    atomicExch(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__double_as_longlong(v));
  } else {
    wv_publish(p, v);
    if (mode == 2) __threadfence();
  }
}
__device__ __forceinline__ void wv_publish_mode(cplx* p, cplx v, int mode) {
  if (mode == 1) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    atomicExch(q + 1, (unsigned long long)__double_as_longlong(v.im));
    atomicExch(q, (unsigned long long)__double_as_longlong(v.re));
  } else {
    wv_publish(p, v);
    if (mode == 2) __threadfence();
  }
}

template <typename T>
struct WaveArgs {
  const unsigned char* stat;
  const WaveChunk* chunks;
  const int* blk_chunk;
  const T* rhsp;
  const T* xg;  // null: the other triangle is skipped (sweep from zero)
  T* out;
  int* ticket;  // [0] block ticket, [1] timeout flag
  int nblocks, block_rows;
  int stages, stage_static, stage_rhs_bytes, stage_bytes;
  long long* stats;  // null, or [4 * nblocks]: clocks total / waiting for the ring / poll retries / thread 0 in the level barrier
  int publish_mode, backoff_ns;
  const int* gate;
  int gate_value;
};

// Pre-pass (fully parallel): sentinel-fill out, bring rhs and the other-side x entries into the
// order the sweep consumes them, reset the block ticket.
template <typename T>
__global__ void __launch_bounds__(kVecThreads) gs_wave_prep_kernel(int64_t n8, unsigned long long* out8, int64_t rhs_slots, const int* rowmap,
                                                                    const T* rhs, T* rhsp, int64_t xg_slots, const int* ocol, const T* other,
                                                                    T* xg, int* ticket, const int* gate, int gate_value) {
  if (gate && *gate != gate_value) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) ticket[0] = 0;
  SPB_GRID_STRIDE(i, n8) out8[i] = SPB_GS_SENTINEL;
  SPB_GRID_STRIDE(i, rhs_slots) {
    const int r = rowmap[i];
    rhsp[i] = r >= 0 ? rhs[r] : zero_of<T>();
  }
  if (other) {
    SPB_GRID_STRIDE(i, xg_slots) {
      const int c = ocol[i];
      xg[i] = c >= 0 ? other[c] : zero_of<T>();
    }
  }
}

// Folds the produced-side entries [k0, k1) of a row into sigma, in order.  Intra-block columns are
// shared-memory reads, the others are polled in global memory until they are values.
template <typename T>
__device__ __forceinline__ T wave_fold_produced(const WaveArgs<T>& a, const int* ecol, const T* eval, const T* xs, int k0, int k1, T sigma,
                                                long long& retries) {
  for (int kb = k0; kb < k1; kb += WAVE_PF) {
    T xv[WAVE_PF];
    unsigned c[WAVE_PF];
    unsigned pending = 0;
#pragma unroll
    for (int j = 0; j < WAVE_PF; ++j) {
      const int k = kb + j;
      c[j] = k < k1 ? (unsigned)ecol[k] : SPB_WAVE_INTRA;
      xv[j] = zero_of<T>();
      if (k < k1) {
        if (c[j] & SPB_WAVE_INTRA) {
          xv[j] = xs[c[j] & ~SPB_WAVE_INTRA];
        } else {
          xv[j] = wv_poll(a.out + c[j]);
          if (wv_is_sentinel(xv[j])) pending |= 1u << j;
        }
      }
    }
    long long spins = 0;
    while (pending) {  // values of other blocks still in flight
      if (a.backoff_ns > 0) __nanosleep(a.backoff_ns);
#pragma unroll
      for (int j = 0; j < WAVE_PF; ++j) {
        if (pending & (1u << j)) {
          xv[j] = wv_poll(a.out + c[j]);
          if (!wv_is_sentinel(xv[j])) pending &= ~(1u << j);
        }
      }
      if (++spins > (1LL << 22)) {  // a legitimate value that equals the sentinel (a NaN): use it
        a.ticket[1] = 1;
        break;
      }
    }
    retries += spins;
#pragma unroll
    for (int j = 0; j < WAVE_PF; ++j)
      if (kb + j < k1) sigma = add(sigma, mul(eval[kb + j], xv[j]));  // src/gauss_seidel.rs:113-118
  }
  return sigma;
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
  const int tid = threadIdx.x;
  const long long t_start = a.stats ? clock64() : 0;
  if (tid == 0) {
    *s_ticket = atomicAdd(a.ticket, 1);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], WAVE_NC / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int t = *s_ticket;  // blocks are processed in ticket order: a block waits only for earlier tickets
  const int b = BWD ? a.nblocks - 1 - t : t;
  const int r0 = b * a.block_rows;
  const int c0 = a.blk_chunk[t], c1 = a.blk_chunk[t + 1];
  const int S = a.stages;

  if (tid >= WAVE_NC) {  // ---- producer warp: one lane streams the block's chunks through the ring
    if (tid == WAVE_NC) {
      const uint64_t pol = l2_policy_evict_first();
      WaveChunk dn = c0 < c1 ? a.chunks[c0] : WaveChunk{};
      for (int c = c0; c < c1; ++c) {
        const int k = c - c0, s = k % S, u = k / S;
        const WaveChunk d = dn;
        if (c + 1 < c1) dn = a.chunks[c + 1];  // descriptor of the next chunk: in flight during the wait
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1));
        unsigned char* st = ring + (size_t)s * a.stage_bytes;
        const uint32_t rb = (uint32_t)wave_a16((long long)sizeof(T) * d.nrows);
        const uint32_t xb = a.xg ? (uint32_t)wave_a16((long long)sizeof(T) * d.xg_cnt) : 0u;
        mbar_arrive_expect_tx(&full[s], (uint32_t)d.sbytes + rb + xb);
        bulk_g2s(st, a.stat + d.soff, (uint32_t)d.sbytes, &full[s], pol);
        bulk_g2s(st + a.stage_static, a.rhsp + d.rhs_off, rb, &full[s], pol);
        if (xb) bulk_g2s(st + a.stage_static + a.stage_rhs_bytes, a.xg + d.xg_off, xb, &full[s], pol);
      }
    }
    return;
  }

  // ---- consumers
  long long wait_clk = 0, retries = 0, bar_clk = 0;
  for (int c = c0; c < c1; ++c) {
    const int k = c - c0, s = k % S, u = k / S;
    if (a.stats && tid == 0) {
      const long long w0 = clock64();
      mbar_wait(&full[s], (uint32_t)(u & 1));
      wait_clk += clock64() - w0;
    } else {
      mbar_wait(&full[s], (uint32_t)(u & 1));
    }
    const unsigned char* st = ring + (size_t)s * a.stage_bytes;
    const int* hdr = reinterpret_cast<const int*>(st);
    const int nrows = hdr[0], nseg = hdr[1], nent = hdr[2];
    const WaveLayout L = wave_layout<T>(nrows, nseg, nent);
    const int* seg_end = reinterpret_cast<const int*>(st + L.seg_end);
    const int* rowid = reinterpret_cast<const int*>(st + L.rowid);
    const int* pptr = reinterpret_cast<const int*>(st + L.pptr);
    const int* nprod = reinterpret_cast<const int*>(st + L.nprod);
    const int* optr = reinterpret_cast<const int*>(st + L.optr);
    const T* diag = reinterpret_cast<const T*>(st + L.diag);
    const T* eval = reinterpret_cast<const T*>(st + L.eval);
    const int* ecol = reinterpret_cast<const int*>(st + L.ecol);
    const T* rhss = reinterpret_cast<const T*>(st + a.stage_static);
    const T* xgs = reinterpret_cast<const T*>(st + a.stage_static + a.stage_rhs_bytes);
    int sbeg = 0;
    for (int g = 0; g < nseg; ++g) {
      const int send = seg_end[g];
      for (int i = sbeg + tid; i < send; i += WAVE_NC) {
        const int p0 = pptr[i], p1 = pptr[i + 1], np = nprod[i];
        T sigma = zero_of<T>();
        if (!BWD) {  // CSR order: lower (produced) entries, then upper (other side)
          sigma = wave_fold_produced<T>(a, ecol, eval, xs, p0, p0 + np, sigma, retries);
          if (a.xg) {
            const T* xo = xgs + optr[i] - (p0 + np);
            for (int q = p0 + np; q < p1; ++q) sigma = add(sigma, mul(eval[q], xo[q]));
          }
        } else {     // lower (other side) entries, then upper (produced)
          if (a.xg) {
            const T* xo = xgs + optr[i] - p0;
            for (int q = p0; q < p1 - np; ++q) sigma = add(sigma, mul(eval[q], xo[q]));
          }
          sigma = wave_fold_produced<T>(a, ecol, eval, xs, p1 - np, p1, sigma, retries);
        }
        const T x = divi(sub(rhss[i], sigma), diag[i]);  // src/gauss_seidel.rs:123
        const int row = rowid[i];
        xs[row - r0] = x;
        wv_publish_mode(a.out + row, x, a.publish_mode);
      }
      const long long w0 = a.stats ? clock64() : 0;
      consumer_bar_sync(WAVE_NC);  // the level is complete: its x values are visible in xs
      if (a.stats) bar_clk += clock64() - w0;
      sbeg = send;
    }
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  if (a.stats) {
    if (retries) atomicAdd(reinterpret_cast<unsigned long long*>(a.stats + 4 * t + 2), (unsigned long long)retries);
    if (tid == 0) {
      a.stats[4 * t + 0] = clock64() - t_start;
      a.stats[4 * t + 1] = wait_clk;
      a.stats[4 * t + 3] = bar_clk;
    }
  }
}

// ---- analysis ----------------------------------------------------------------------------------
template <typename T>
static void put_bytes(std::vector<unsigned char>& v, size_t off, const T* src, size_t count) {
  if (count) memcpy(v.data() + off, src, sizeof(T) * count);
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
  ws.stage_other = std::max(64, env("SPB_GS_STAGE_OTHER", 1024));
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
  ws.smem_bytes = (size_t)WAVE_FIXED + ((size_t)R * sizeof(T) + 127) / 128 * 128 + (size_t)ws.stages * stage_bytes;

  const size_t align_el = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;  // elements per 16 bytes
  auto align_slots = [&](int64_t v) { return (int64_t)((v + align_el - 1) / align_el * align_el); };

  std::vector<int> lev(n, 0);
  std::vector<unsigned char> stat;
  std::vector<WaveChunk> chunks;
  std::vector<int> blk_chunk(nb + 1, 0), rowmap, ocol;
  stat.reserve((size_t)(ip[n] * (sizeof(T) + 4) + n * (24 + sizeof(T))));
  rowmap.reserve((size_t)n + 2 * (size_t)nb);
  int64_t max_levels = 0;

  std::vector<int> order, cnt;  // rows of the block sorted by (local level, row)
  // open chunk
  std::vector<int> ch_rows, ch_segend;
  int ch_nent = 0, ch_noth = 0;
  bool seg_open = false;

  auto row_counts = [&](int64_t i, int& ne, int& no, bool& sorted_ok) {
    ne = no = 0;
    bool seen_upper = false;
    for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
      const int j = cols[k];
      if (j == i) continue;
      if (j < 0 || j >= n) sorted_ok = false;
      if (j < i) {
        if (seen_upper) sorted_ok = false;  // a lower entry after an upper one: the CSR-order fold would change
      } else {
        seen_upper = true;
      }
      ++ne;
      if (backward ? (j < i) : (j > i)) ++no;
    }
  };

  auto close_chunk = [&](int64_t r0, int64_t r1) {
    if (ch_rows.empty()) return;
    if (seg_open) ch_segend.push_back((int)ch_rows.size());
    seg_open = false;
    const int nrows = (int)ch_rows.size(), nseg = (int)ch_segend.size();
    const WaveLayout L = wave_layout<T>(nrows, nseg, ch_nent);
    const size_t base = stat.size();
    stat.resize(base + L.total, 0);
    WaveChunk d{};
    d.soff = (long long)base;
    d.sbytes = L.total;
    d.nrows = nrows;
    d.rhs_off = (long long)rowmap.size();
    d.xg_off = (long long)ocol.size();
    std::vector<int> pptr(nrows + 1), nprod(nrows), optr(nrows), ecol(ch_nent);
    std::vector<T> dg(nrows), ev(ch_nent);
    int e = 0, o = 0;
    for (int q = 0; q < nrows; ++q) {
      const int64_t i = ch_rows[q];
      pptr[q] = e;
      optr[q] = o;
      int np = 0;
      T dv = zero_of<T>();
      for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
        const int j = cols[k];
        if (j == i) {
          dv = vals[k];  // the last diagonal entry wins, as in the loop of src/gauss_seidel.rs:119-121
          continue;
        }
        const bool produced = backward ? (j > i) : (j < i);
        ev[e] = vals[k];
        if (produced) {
          ecol[e] = (j >= r0 && j < r1) ? (int)(SPB_WAVE_INTRA | (unsigned)(j - r0)) : j;
          ++np;
        } else {
          ecol[e] = j;
          ocol.push_back(j);
          ++o;
        }
        ++e;
      }
      nprod[q] = np;
      dg[q] = dv;
      rowmap.push_back((int)i);
    }
    pptr[nrows] = e;
    d.xg_cnt = o;
    while ((int64_t)rowmap.size() != align_slots((int64_t)rowmap.size())) rowmap.push_back(-1);
    while ((int64_t)ocol.size() != align_slots((int64_t)ocol.size())) ocol.push_back(-1);
    const int hdr[4] = {nrows, nseg, ch_nent, 0};
    put_bytes(stat, base, hdr, 4);
    put_bytes(stat, base + L.seg_end, ch_segend.data(), nseg);
    put_bytes(stat, base + L.rowid, ch_rows.data(), nrows);
    put_bytes(stat, base + L.pptr, pptr.data(), nrows + 1);
    put_bytes(stat, base + L.nprod, nprod.data(), nrows);
    put_bytes(stat, base + L.optr, optr.data(), nrows);
    put_bytes(stat, base + L.diag, dg.data(), nrows);
    put_bytes(stat, base + L.eval, ev.data(), ch_nent);
    put_bytes(stat, base + L.ecol, ecol.data(), ch_nent);
    chunks.push_back(d);
    ch_rows.clear();
    ch_segend.clear();
    ch_nent = ch_noth = 0;
  };

  // GLOBAL levels (longest dependency chain over the whole triangle).  Every block walks its rows
  // in global-level order, so all blocks advance along the same wavefront: a value another block
  // needs at level l was produced at a level < l, i.e. earlier on the producer's own time line,
  // whatever the block boundaries are.  (Ordering by block-local levels serialises the blocks
  // whenever the boundaries do not line up with the structure of the matrix.)
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
    for (int l = 0; l < nl; ++l) {
      for (int q = cnt[l]; q < cnt[l + 1]; ++q) {
        const int64_t i = order[q];
        int ne, no;
        bool sorted_ok = true;
        row_counts(i, ne, no, sorted_ok);
        if (!sorted_ok) return;
        if (wave_layout<T>(1, 1, ne).total > ws.stage_static || no > ws.stage_other) return;  // row too long for a stage
        const int nseg_new = (int)ch_segend.size() + 1;
        if (!ch_rows.empty() &&
            ((int)ch_rows.size() + 1 > ws.stage_rows || ch_noth + no > ws.stage_other ||
             wave_layout<T>((int)ch_rows.size() + 1, nseg_new, ch_nent + ne).total > ws.stage_static))
          close_chunk(r0, r1);
        ch_rows.push_back((int)i);
        ch_nent += ne;
        ch_noth += no;
        seg_open = true;
      }
      if (seg_open) {  // end of a local level: barrier point
        ch_segend.push_back((int)ch_rows.size());
        seg_open = false;
      }
    }
    close_chunk(r0, r1);
    if (chunks.size() >= (size_t)1 << 31) return;
  }
  blk_chunk[nb] = (int)chunks.size();
  ws.nchunks = (int64_t)chunks.size();
  ws.rhs_slots = (int64_t)rowmap.size();
  ws.xg_slots = (int64_t)ocol.size();
  ws.local_levels_max = max_levels;

  ws.stat.alloc(stat.size() + 64);
  ws.chunks.alloc(sizeof(WaveChunk) * std::max<size_t>(chunks.size(), 1));
  ws.blk_chunk.alloc(sizeof(int) * (nb + 1));
  ws.rowmap.alloc(sizeof(int) * std::max<size_t>(rowmap.size(), 1));
  ws.ocol.alloc(sizeof(int) * std::max<size_t>(ocol.size(), 1));
  ws.rhsp.alloc(sizeof(T) * (rowmap.size() + 4));
  ws.xg.alloc(sizeof(T) * (ocol.size() + 4));
  ws.ticket.alloc(sizeof(int) * 4);
  SPB_CUDA(cudaMemcpyAsync(ws.stat.p, stat.data(), stat.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.chunks.p, chunks.data(), sizeof(WaveChunk) * chunks.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.blk_chunk.p, blk_chunk.data(), sizeof(int) * (nb + 1), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemcpyAsync(ws.rowmap.p, rowmap.data(), sizeof(int) * rowmap.size(), cudaMemcpyHostToDevice, c->stream));
  if (!ocol.empty()) SPB_CUDA(cudaMemcpyAsync(ws.ocol.p, ocol.data(), sizeof(int) * ocol.size(), cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaMemsetAsync(ws.ticket.p, 0, sizeof(int) * 4, c->stream));
  SPB_CUDA(cudaMemsetAsync(ws.rhsp.p, 0, ws.rhsp.bytes, c->stream));
  SPB_CUDA(cudaMemsetAsync(ws.xg.p, 0, ws.xg.bytes, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
  ws.ok = true;
}

template <typename T>
void wave_sweep(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, T* out) {
  Ctx* c = M->ctx;
  const int64_t n = M->A->n_local;
  if (n == 0) return;
  const int rhs_bytes = wave_a16((long long)sizeof(T) * ws.stage_rows);
  const int oth_bytes = wave_a16((long long)sizeof(T) * ws.stage_other);
  const int stage_bytes = (ws.stage_static + rhs_bytes + oth_bytes + 127) / 128 * 128;
  {
    LaunchScope lsc(c, FAM_PRECOND);
    const int64_t n8 = n * (int64_t)(sizeof(T) / 8);
    const int64_t work = std::max(n8, std::max(ws.rhs_slots, other ? ws.xg_slots : (int64_t)0));
    gs_wave_prep_kernel<T><<<vec_grid(c, work), kVecThreads, 0, c->stream>>>(
        n8, reinterpret_cast<unsigned long long*>(out), ws.rhs_slots, bufptr<int>(ws.rowmap), rhs, bufptr<T>(ws.rhsp), ws.xg_slots,
        bufptr<int>(ws.ocol), other, bufptr<T>(ws.xg), bufptr<int>(ws.ticket), c->gate, c->gate_value);
    check_launch("gs_wave_prep_kernel");
  }
  WaveArgs<T> a{};
  a.stat = bufptr<unsigned char>(ws.stat);
  a.chunks = bufptr<WaveChunk>(ws.chunks);
  a.blk_chunk = bufptr<int>(ws.blk_chunk);
  a.rhsp = bufptr<T>(ws.rhsp);
  a.xg = other ? bufptr<T>(ws.xg) : nullptr;
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
  static int publish_mode = -1, backoff_ns = 0;
  if (publish_mode < 0) {
    const char* e = getenv("SPB_GS_PUBLISH");
    publish_mode = e && *e ? atoi(e) : 0;
    e = getenv("SPB_GS_BACKOFF");
    backoff_ns = e && *e ? atoi(e) : 0;
  }
  a.publish_mode = publish_mode;
  a.backoff_ns = backoff_ns;
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
