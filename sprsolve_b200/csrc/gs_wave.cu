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
//   * a dependency on another block travels through a MAILBOX in global memory: at analysis time
//     every chunk of a consumer block gets a contiguous range of mailbox slots, one per distinct
//     outside column, in the order it needs them, and the producing row gets the list of slots
//     that want its value.  The mailbox is pre-filled with a sentinel (a NaN payload no arithmetic
//     produces) and a value is delivered with ONE 8/16-byte store per slot, so the value itself
//     says that it is ready -- no flags, no fences.  The polling is done by four HELPER warps per
//     CTA that run ahead of the row threads; their loads are COALESCED (consecutive slots), which
//     matters because a hand-off is paid in the polling SM's own load path (841 clocks unloaded,
//     ~2000 with 127 scattered polls in flight, tools/micro/hop_bench.cu).  They drop each value
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
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include "ops.cuh"
#include "tma.cuh"
#include "vecops.cuh"

namespace spb {

#ifndef SPB_WAVE_NC
#define SPB_WAVE_NC 256
#endif
static const int WAVE_NC = SPB_WAVE_NC;                  // row threads
static const int WAVE_NH = 128;                          // helper threads (cross-block values)
static const int WAVE_HB = 4;                            // polls a helper thread keeps in flight
static const int WAVE_THREADS = WAVE_NC + WAVE_NH + 32;  // + one producer warp
static const int WAVE_MAX_STAGES = 8;
static const int WAVE_FIXED = 256;     // barriers, ticket slot, zero slot
static const int WAVE_ZERO_OFF = 192;  // 16 bytes of +0.0: target of the ELL padding entries
#define SPB_GS_SENTINEL 0xFFFFDEADBEEF5EEDULL
#define SPB_GS_SENTINEL32 0xFFDEAD5Eu /* f32 / Complex32 slots: a NaN payload in every 32-bit word */
// the 64-bit word the pre-pass fills the mailbox with: one f64 sentinel or two f32 sentinels
template <typename T>
__host__ __device__ inline unsigned long long wave_sentinel_fill() {
  return sizeof(real_t<T>) == 4 ? (((unsigned long long)SPB_GS_SENTINEL32 << 32) | SPB_GS_SENTINEL32) : SPB_GS_SENTINEL;
}

__host__ __device__ inline int wave_a16(long long v) { return (int)((v + 15) & ~15LL); }

// Byte offsets of the sections of one packed chunk (all 16-byte aligned).
// header (64 B): nrows, nseg, W, nhalo, Wo, Wm, mailbox offset (lo, hi); then the byte offsets of the sections
// (seg_end, rowid, diag, eoff, eval, mail, hslot, hcol)
#define SPB_WAVE_NOMAIL 0xFFFFFFFFu
struct WaveLayout {
  int seg_end, rowid, diag, eoff, eval, mail, hslot, hcol, total;
};
template <typename T>
__host__ __device__ inline WaveLayout wave_layout(int nrows, int nseg, int W, int nhalo, int Wm) {
  WaveLayout L;
  int off = 64;  // header: 8 ints of sizes + the 8 section offsets below (the kernel reads them instead of recomputing)
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
  L.mail = off;  // ELL, column-major: mailbox slots (of later blocks) that want the row's value
  off += wave_a16(4LL * Wm * nrows);
  L.hslot = off;  // sentinel-filled landing slots of the cross-block values (mailbox order)
  off += wave_a16((long long)sizeof(T) * nhalo);
  L.hcol = off;  // producer of every landing slot, block << 16 | row in block (cluster mode: whose shared memory holds it)
  off += wave_a16(4LL * nhalo);
  L.total = off;
  return L;
}

// the layout of a staged chunk as the analysis stored it in the header (two 16-byte shared-memory loads)
__device__ __forceinline__ WaveLayout wave_layout_of(const int* hdr) {
  const int4 a = *reinterpret_cast<const int4*>(hdr + 8), b = *reinterpret_cast<const int4*>(hdr + 12);
  WaveLayout L;
  L.seg_end = a.x;
  L.rowid = a.y;
  L.diag = a.z;
  L.eoff = a.w;
  L.eval = b.x;
  L.mail = b.y;
  L.hslot = b.z;
  L.hcol = b.w;
  L.total = 0;
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
__device__ __forceinline__ float wv_poll(const float* p) {
  float v;
  asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ cplxf wv_poll(const cplxf* p) {
  cplxf v;
  asm volatile("ld.relaxed.gpu.global.v2.f32 {%0,%1}, [%2];" : "=f"(v.re), "=f"(v.im) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wv_is_sentinel(double v) {
  return (unsigned long long)__double_as_longlong(v) == SPB_GS_SENTINEL;
}
__device__ __forceinline__ bool wv_is_sentinel(cplx v) { return wv_is_sentinel(v.re) || wv_is_sentinel(v.im); }
__device__ __forceinline__ bool wv_is_sentinel(float v) { return (unsigned)__float_as_int(v) == SPB_GS_SENTINEL32; }
__device__ __forceinline__ bool wv_is_sentinel(cplxf v) { return wv_is_sentinel(v.re) || wv_is_sentinel(v.im); }
// A computed x that happens to BE the sentinel pattern (only possible when the caller's rhs carries a NaN with
// exactly this payload: NaN payloads propagate through the arithmetic) is handed on as a different NaN, so a
// consumer never mistakes it for "not there yet"; the output vector keeps the original bits.
__device__ __forceinline__ double wv_handoff(double v) {
  return wv_is_sentinel(v) ? __longlong_as_double((long long)(SPB_GS_SENTINEL ^ 1ULL)) : v;
}
__device__ __forceinline__ float wv_handoff(float v) { return wv_is_sentinel(v) ? __int_as_float((int)(SPB_GS_SENTINEL32 ^ 1u)) : v; }
__device__ __forceinline__ cplx wv_handoff(cplx v) { return cplx{wv_handoff(v.re), wv_handoff(v.im)}; }
__device__ __forceinline__ cplxf wv_handoff(cplxf v) { return cplxf{wv_handoff(v.re), wv_handoff(v.im)}; }
__device__ __forceinline__ void wv_publish(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void wv_publish(cplx* p, cplx v) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.re), "d"(v.im) : "memory");
}
__device__ __forceinline__ void wv_publish(float* p, float v) {
  asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void wv_publish(cplxf* p, cplxf v) {
  asm volatile("st.relaxed.gpu.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.re), "f"(v.im) : "memory");
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
template <>
__device__ __forceinline__ float wv_lds<float>(uint32_t addr) {
  float v;
  asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
template <>
__device__ __forceinline__ cplxf wv_lds<cplxf>(uint32_t addr) {
  cplxf v;
  asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.re), "=f"(v.im) : "r"(addr) : "memory");
  return v;
}
// ---- thread-block cluster: the x block of a neighbouring CTA of the same cluster is polled in ITS shared memory
__device__ __forceinline__ uint32_t wv_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t wv_cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void wv_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t wv_mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ int wv_ld_remote_s32(uint32_t ra) {
  int v;
  asm volatile("ld.volatile.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
  return v;
}
template <typename T>
__device__ __forceinline__ T wv_poll_remote(uint32_t ra);
template <>
__device__ __forceinline__ double wv_poll_remote<double>(uint32_t ra) {
  double v;
  asm volatile("ld.volatile.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
template <>
__device__ __forceinline__ cplx wv_poll_remote<cplx>(uint32_t ra) {
  cplx v;
  asm volatile("ld.volatile.shared::cluster.v2.f64 {%0,%1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(ra) : "memory");
  return v;
}
template <>
__device__ __forceinline__ float wv_poll_remote<float>(uint32_t ra) {
  float v;
  asm volatile("ld.volatile.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
template <>
__device__ __forceinline__ cplxf wv_poll_remote<cplxf>(uint32_t ra) {
  cplxf v;
  asm volatile("ld.volatile.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.re), "=f"(v.im) : "r"(ra) : "memory");
  return v;
}
template <typename T>
__device__ __forceinline__ void wv_sts_volatile(T* p, T v) {
  *reinterpret_cast<volatile T*>(p) = v;
}
template <>
__device__ __forceinline__ void wv_sts_volatile<cplx>(cplx* p, cplx v) {
  asm volatile("st.volatile.shared.v2.f64 [%0], {%1,%2};" ::"r"(smem_u32(p)), "d"(v.re), "d"(v.im) : "memory");
}
template <>
__device__ __forceinline__ void wv_sts_volatile<cplxf>(cplxf* p, cplxf v) {
  asm volatile("st.volatile.shared.v2.f32 [%0], {%1,%2};" ::"r"(smem_u32(p)), "f"(v.re), "f"(v.im) : "memory");
}
template <typename T>
__device__ __forceinline__ T wave_sentinel_value() {
  const unsigned long long w[2] = {wave_sentinel_fill<T>(), wave_sentinel_fill<T>()};
  T v;
  memcpy(&v, w, sizeof(T));
  return v;
}

// x value at a resolved shared-memory address; spins while the slot still holds the sentinel
// (a cross-block value the helper warps have not delivered yet).
template <typename T>
__device__ __forceinline__ T wv_x(uint32_t addr, int* flag, int* err, long long& spins, unsigned spin_ns) {
  T x = wv_lds<T>(addr);
  if (wv_is_sentinel(x)) {
    long long n = 0;
    do {
      if (spin_ns) __nanosleep(spin_ns);  // leave the SM's load / store path to the helper warps that deliver the value
      x = wv_lds<T>(addr);
    } while (wv_is_sentinel(x) && ++n < (1LL << 26));
    if (wv_is_sentinel(x)) {  // a legitimate value that equals the sentinel (a NaN), or a stalled producer: use it, report it
      flag[0] = 1;
      if (err) *err = 2;
    }
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
  T* mailbox;
  int* ticket;  // [0] block ticket, [1] timeout flag
  int nblocks, block_rows;
  int stages, stage_static, stage_rhs_bytes, stage_bytes;
  long long* stats;  // null, or [4 * nblocks]: clocks total / waiting for the ring / shared-memory spins / thread 0 in the level barrier
  const int* gate;
  int gate_value;
  int* err;  // Ctx::dev_err
  unsigned spin_ns;  // back-off of a row thread that waits for a value of another block
  T* sig;            // forward sweep of the symmetric apply: the fold over the lower entries of every row, for the backward pre-pass
};

// Pre-pass (fully parallel): sentinel-fill the mailbox, permute rhs into sweep order, reduce the other
// triangle to what the sweep needs (see the header), reset the block ticket.
template <typename T, typename IP, bool BWD>
__global__ void __launch_bounds__(kVecThreads) gs_wave_prep_kernel(int64_t mb8, unsigned long long* mailbox8, int64_t rhs_slots, const int* rowmap,
                                                                    const T* rhs, T* rhsp, const T* other, const IP* indptr, const int* cols,
                                                                    const T* vals, T* aux, const long long* aux_base, const int* aux_dims,
                                                                    int* ticket, const int* gate, int gate_value, const T* sig) {
  if (gate && *gate != gate_value) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) ticket[0] = 0;
  SPB_GRID_STRIDE(i, mb8) mailbox8[i] = wave_sentinel_fill<T>();
  SPB_GRID_STRIDE(p, rhs_slots) {
    const int r = rowmap[p];
    rhsp[p] = r >= 0 ? rhs[r] : zero_of<T>();
    if (!other) continue;
    if (BWD && sig) {
      // symmetric apply: `other` is the forward sweep's result, and that sweep has just folded exactly these entries
      // in exactly this order (its own sigma of row r) -- take its value instead of traversing the matrix again
      aux[p] = r >= 0 ? sig[r] : zero_of<T>();
    } else if (BWD) {  // lower entries come first in CSR order: fold them now (src/gauss_seidel.rs:113-118)
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

// CL: the grid is launched in thread-block clusters; block tickets are handed out per cluster (rank r of the cluster
// with ticket c sweeps block c * size + r), so consecutive blocks are co-resident CTAs of one cluster and a value
// produced by one of them is polled by the helpers directly in the producer's shared-memory x block (ld.shared::cluster,
// ~200 clocks) instead of the global mailbox (an L2 round trip per poll: ~3.2 k clocks of lag per block boundary measured).
template <typename T, bool BWD, bool CL>
__global__ void __launch_bounds__(WAVE_THREADS, 1) gs_wave_kernel(const WaveArgs<T> a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (a.gate && *a.gate != a.gate_value) return;  // (uniform over the grid)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + WAVE_MAX_STAGES;
  int* s_ticket = reinterpret_cast<int*>(smem + 2 * WAVE_MAX_STAGES * sizeof(uint64_t));
  T* xs = reinterpret_cast<T*>(smem + WAVE_FIXED);
  unsigned char* ring = smem + WAVE_FIXED + ((size_t)a.block_rows * sizeof(T) + 127) / 128 * 128;
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x;
  const long long t_start = a.stats ? clock64() : 0;
  const uint32_t crank = CL ? wv_cluster_rank() : 0u, csize = CL ? wv_cluster_size() : 1u;
  if (tid == 0) {
    if (crank == 0) *s_ticket = atomicAdd(a.ticket, 1);
    *reinterpret_cast<T*>(smem + WAVE_ZERO_OFF) = zero_of<T>();
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (WAVE_NC + WAVE_NH) / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CL) {  // the x block starts as "nothing produced yet": its neighbours in the cluster poll it
    const T sv = wave_sentinel_value<T>();
    for (int i = tid; i < a.block_rows; i += WAVE_THREADS) xs[i] = sv;
    wv_cluster_sync();
  } else {
    __syncthreads();
  }
  // blocks are processed in ticket order: a block waits only for earlier tickets
  const int t = CL ? wv_ld_remote_s32(wv_mapa(smem_u32(s_ticket), 0)) * (int)csize + (int)crank : *s_ticket;
  if (CL && t >= a.nblocks) {  // padding CTA of the last cluster
    wv_cluster_sync();
    return;
  }
  const int cshift = 31 - __clz((int)csize);  // log2(cluster size)
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
    // (all four warps work on the same chunk: one warp per chunk, the warps taking the chunks round robin, was measured
    //  40 % slower -- 32 lanes with 8 polls each deliver a chunk later than 128 threads with 2)
    const int htid = tid - WAVE_NC;
    for (int c = c0; c < c1; ++c) {
      const int k = c - c0, s = k % S, u = k / S;
      mbar_wait(&full[s], (uint32_t)(u & 1));
      unsigned char* st = ring + (size_t)s * a.stage_bytes;
      const int* hdr = reinterpret_cast<const int*>(st);
      const int nhalo = hdr[3];
      if (nhalo > 0) {
        const WaveLayout L = wave_layout_of(hdr);
        T* hslot = reinterpret_cast<T*>(st + L.hslot);
        const int* hcol = reinterpret_cast<const int*>(st + L.hcol);
        const T* mbox = a.mailbox + (((long long)hdr[7] << 32) | (unsigned)hdr[6]);  // this chunk's slots: contiguous
        // where slot h is polled: the producer's shared-memory x block when it runs in this cluster, else the mailbox
        auto poll = [&](int h, uint32_t ra) -> T { return (CL && ra) ? wv_poll_remote<T>(ra) : wv_poll(mbox + h); };
        auto remote_of = [&](int h) -> uint32_t {  // (cluster sizes are powers of two: shifts, no division)
          if (!CL) return 0u;
          const unsigned w = (unsigned)hcol[h];  // producer block << 16 | row inside that block
          const int pb = (int)(w >> 16);
          const int tp = BWD ? a.nblocks - 1 - pb : pb;
          if ((tp >> cshift) != (t >> cshift)) return 0u;
          return wv_mapa(smem_u32(xs + (w & 0xFFFFu)), (uint32_t)(tp & ((int)csize - 1)));
        };
        // slots are sorted by need; thread t polls slots t, t + NH, ...: coalesced, and every thread starts early
        for (int h0 = htid; h0 < nhalo; h0 += WAVE_NH * WAVE_HB) {
          unsigned pend = 0;
          T v[WAVE_HB];
          uint32_t ra[WAVE_HB];
#pragma unroll
          for (int j = 0; j < WAVE_HB; ++j) {
            ra[j] = 0u;
            if (h0 + j * WAVE_NH < nhalo) {
              ra[j] = remote_of(h0 + j * WAVE_NH);
              v[j] = poll(h0 + j * WAVE_NH, ra[j]);
            }
          }
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
                const T w = poll(h0 + j * WAVE_NH, ra[j]);
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
    const int nrows = hdr[0], nseg = hdr[1], W = hdr[2], Wo = hdr[4], Wm = hdr[5];
    const WaveLayout L = wave_layout_of(hdr);
    const uint32_t* mail = reinterpret_cast<const uint32_t*>(st + L.mail);
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
            x0 = wv_x<T>(sbase + o0, a.ticket + 1, a.err, spins, a.spin_ns);
            x1 = wv_x<T>(sbase + o1, a.ticket + 1, a.err, spins, a.spin_ns);
            x2 = wv_x<T>(sbase + o2, a.ticket + 1, a.err, spins, a.spin_ns);
            x3 = wv_x<T>(sbase + o3, a.ticket + 1, a.err, spins, a.spin_ns);
          }
          sigma = add(sigma, mul(v0, x0));
          sigma = add(sigma, mul(v1, x1));
          sigma = add(sigma, mul(v2, x2));
          sigma = add(sigma, mul(v3, x3));
        }
        if (!BWD && a.sig) a.sig[row] = sigma;  // (the backward sweep of the symmetric apply starts from this fold)
        if (!BWD && a.aux)
          for (int e = 0; e < Wo; ++e) sigma = add(sigma, auxs[e * nrows + i]);
        const T x = divi(sub(rv, sigma), dv);  // src/gauss_seidel.rs:123
        const T xh = wv_handoff(x);
        if (CL)
          wv_sts_volatile(xs + (row - r0), xh);  // polled by the next blocks of the cluster
        else
          xs[row - r0] = xh;
        for (int e = 0; e < Wm; ++e) {  // deliver to the blocks that wait for this value
          const uint32_t m = mail[e * nrows + i];
          if (m != SPB_WAVE_NOMAIL) wv_publish(a.mailbox + m, xh);
        }
        a.out[row] = x;
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
  if (CL) wv_cluster_sync();  // the x block stays addressable until every block of the cluster is done (row warps only:
                              // the producer / helper warps have exited or will, and exited threads are not waited for)
}

// ---- analysis ----------------------------------------------------------------------------------
template <typename U>
static void put_bytes(std::vector<unsigned char>& v, size_t off, const U* src, size_t count) {
  if (count) memcpy(v.data() + off, src, sizeof(U) * count);
}

// Far stride of the produced triangle: per sampled row, the median of the cluster of produced-side offsets within
// 10 % of the farthest one (7-point: {plane}; 27-point: the nine entries of the neighbouring plane -> its centre);
// returned when at least 60 % of the sampled rows agree (+-1), else 0 (no grid structure).
static int64_t wave_far_stride(const std::vector<int64_t>& ip, const std::vector<int>& cols, int64_t n, bool backward) {
  std::vector<int64_t> far;
  const int64_t step = std::max<int64_t>(1, n / 4096);
  std::vector<int64_t> offs;
  for (int64_t r = 0; r < n; r += step) {
    offs.clear();
    for (int64_t k = ip[r]; k < ip[r + 1]; ++k) {
      const int64_t d = backward ? (int64_t)cols[k] - r : r - (int64_t)cols[k];
      if (d > 0) offs.push_back(d);
    }
    if (offs.empty()) continue;
    std::sort(offs.begin(), offs.end());
    const int64_t dmax = offs.back();
    size_t lo = offs.size() - 1;
    while (lo > 0 && offs[lo - 1] * 10 >= dmax * 9) --lo;
    far.push_back(offs[lo + (offs.size() - lo) / 2]);
  }
  if (far.size() < 8) return 0;
  std::vector<int64_t> srt = far;
  std::nth_element(srt.begin(), srt.begin() + srt.size() / 2, srt.end());
  const int64_t med = srt[srt.size() / 2];
  size_t agree = 0;
  for (int64_t v : far) agree += (v >= med - 1 && v <= med + 1);
  return agree * 10 >= far.size() * 6 ? med : 0;
}

template <typename T>
void wave_build(CsrMat<T>* A, const std::vector<int64_t>& ip, const std::vector<int>& cols, const std::vector<T>& vals,
                bool backward, WaveSched& ws) {
  Ctx* c = A->ctx;
  const int64_t n = A->n_local;
  ws.ok = false;
  ws.backward = backward;
  const bool timing = getenv("SPB_GS_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gs_wave %s] %-12s %.1f ms\n", backward ? "bwd" : "fwd", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  if (n <= 0 || n >= ((int64_t)1 << 31) - 1 || getenv("SPB_GS_LEGACY")) return;
  auto env = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
  };
  ws.stages = std::max(2, std::min(WAVE_MAX_STAGES, env("SPB_GS_STAGES", 3)));
  // Stage capacity: 16 KB holds ~256 rows of a 7-point matrix.  Fat rows (27-point: ~240 B) would
  // leave only one or two levels per chunk and the per-chunk costs (ring hand-shake, one poll round
  // trip of the helpers) would dominate: give them 32 KB when the x block still fits next to it.
  int dflt_static = 16384, dflt_other = 1024;
  {
    const double row_bytes = 12.0 + sizeof(T) + (double)(ip[n] - n) / (double)n * 0.5 * (4.0 + sizeof(T)) * 1.15 + 16.0;
    const int64_t need_rows = ceil_div(n, c->sm_count);
    int smem_cap = 0;
    SPB_CUDA(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const int64_t big_stage = 32768 + wave_a16((long long)sizeof(T) * 256) + wave_a16((long long)sizeof(T) * 2048) + 128;
    if (row_bytes > 110.0 && WAVE_FIXED + 512 + 3 * big_stage + need_rows * (int64_t)sizeof(T) <= smem_cap) {
      dflt_static = 32768;
      dflt_other = 2048;
    }
  }
  ws.stage_static = wave_a16(std::max(1024, env("SPB_GS_STAGE_BYTES", dflt_static)));
  ws.stage_rows = std::max(8, env("SPB_GS_STAGE_ROWS", 256));
  ws.stage_other = std::max(ws.stage_rows, env("SPB_GS_STAGE_OTHER", dflt_other));
  const int rhs_bytes = wave_a16((long long)sizeof(T) * ws.stage_rows);
  const int oth_bytes = wave_a16((long long)sizeof(T) * ws.stage_other);
  const int stage_bytes = (ws.stage_static + rhs_bytes + oth_bytes + 127) / 128 * 128;
  int smem_max = 0;
  SPB_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  const int64_t avail = (int64_t)smem_max - WAVE_FIXED - (int64_t)ws.stages * stage_bytes - 256;
  const int64_t rmax = avail / (int64_t)sizeof(T);
  if (rmax < 64) return;
  int64_t R = env("SPB_GS_BLOCK_ROWS", 0);
  if (R <= 0) {
    R = std::max<int64_t>(256, ceil_div(n, c->sm_count));
    // Grid-like matrices: cut the blocks at multiples of the far stride of the produced triangle (the plane of a
    // 3-D stencil, the line of a 2-D one), so only the far dependency crosses blocks and every block meets the
    // wavefront at the same relative rows.  Measured on B200 (profiles/r02_gs_block_rows.txt): 7-point 128^3 SGS
    // apply 0.928 -> 0.808 ms, 100^3 0.678 -> 0.573 ms; 27-point 96^3 2.226 -> 2.202 ms.
    const int64_t S = wave_far_stride(ip, cols, n, backward);
    if (S > 1 && !getenv("SPB_GS_NO_STRIDE")) {
      int64_t m = std::max<int64_t>(1, (2 * R + S) / (2 * S));  // round(R / S)
      while (m > 1 && m * S > rmax) --m;
      int64_t Rs = m * S;
      if (Rs > rmax) Rs = ceil_div(S, ceil_div(S, rmax));  // an integer fraction of the stride
      if (Rs >= 64) R = Rs;
    }
  }
  R = std::max<int64_t>(1, std::min(R, rmax));
  const int64_t nb = ceil_div(n, R);
  ws.block_rows = (int)R;
  ws.nblocks = (int)nb;
  if (R > 65535 || nb > 65535) ws.cluster = 0;  // (the packed producer words of the cluster mode hold 16 bits each)
  const size_t ring_off = (size_t)WAVE_FIXED + ((size_t)R * sizeof(T) + 127) / 128 * 128;
  ws.smem_bytes = ring_off + (size_t)ws.stages * stage_bytes;

  const int64_t align_el = std::max<int64_t>(1, 16 / (int64_t)sizeof(T));  // elements per 16 bytes
  auto align_slots = [&](int64_t v) { return (v + align_el - 1) / align_el * align_el; };
  auto pad4 = [](int v) { return (v + 3) & ~3; };
  const unsigned long long sentinel = wave_sentinel_fill<T>();

  std::vector<int> lev(n, 0);
  std::vector<unsigned char> stat;
  std::vector<WaveChunk> chunks;
  std::vector<int> blk_chunk(nb + 1, 0), rowmap, aux_dims;
  std::vector<long long> aux_base;
  stat.reserve((size_t)(ip[n] * (sizeof(T) + 4) + n * (24 + sizeof(T))));
  rowmap.reserve((size_t)n + 2 * (size_t)nb);
  int64_t max_levels = 0, aux_slots = 0;

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
    int gl = 0;
    for (int64_t i = 0; i < n; ++i) gl = std::max(gl, lev[i]);
    ws.global_levels = (int64_t)gl + 1;
  }

  lap("levels");
  // per-row counts: produced entries, other-side entries, produced entries outside the block;
  // ext_out[j]: how many rows of OTHER blocks read x_j (an upper bound of the mailbox slots row j
  // delivers to -- slots are shared by the rows of one consumer chunk)
  struct RowInfo {
    int np, no, nx;
  };
  std::vector<RowInfo> rinfo(n);
  std::vector<int> ext_out(n, 0);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r0 = (i / R) * R, r1 = std::min(n, r0 + R);
    RowInfo ri{0, 0, 0};
    bool seen_upper = false;
    for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
      const int j = cols[k];
      if (j == i) continue;
      if (j < 0 || j >= n) return;
      if (j < i) {
        if (seen_upper) return;  // a lower entry after an upper one: the CSR-order fold would change
      } else {
        seen_upper = true;
      }
      if (backward ? (j > i) : (j < i)) {
        ++ri.np;
        if (j < r0 || j >= r1) {
          ++ri.nx;
          ++ext_out[j];
        }
      } else {
        ++ri.no;
      }
    }
    rinfo[i] = ri;
  }

  lap("row info");
  // ---- pass A: chunk membership, mailbox ranges ------------------------------------------------
  struct Plan {
    int row_beg, row_end;  // into plan_rows
    int seg_beg, seg_end;  // into plan_segs
    int hal_beg, hal_end;  // into plan_halo (distinct outside columns, in need order)
    int W, Wo, kblk;
    long long mb_off;
  };
  std::vector<Plan> plans;
  std::vector<int> plan_rows, plan_segs, plan_halo, order, cnt;
  plan_rows.reserve((size_t)n);
  std::vector<int> seen(n, -1);  // seen[j] == chunk id: column j already has a slot in the open chunk
  long long mb_total = 0;

  int ch_row_beg = 0, ch_seg_beg = 0, ch_hal_beg = 0, ch_w = 0, ch_wo = 0, ch_wm_est = 0;
  bool seg_open = false;
  auto chunk_rows = [&]() { return (int)plan_rows.size() - ch_row_beg; };
  auto chunk_halo = [&]() { return (int)plan_halo.size() - ch_hal_beg; };
  auto close_chunk = [&](int kblk) {
    if (chunk_rows() == 0) return;
    if (seg_open) plan_segs.push_back(chunk_rows());
    seg_open = false;
    Plan pl{ch_row_beg, (int)plan_rows.size(), ch_seg_beg, (int)plan_segs.size(), ch_hal_beg, (int)plan_halo.size(),
            pad4(ch_w), backward ? 0 : ch_wo, kblk, mb_total};
    mb_total = align_slots(mb_total + chunk_halo());
    plans.push_back(pl);
    ch_row_beg = (int)plan_rows.size();
    ch_seg_beg = (int)plan_segs.size();
    ch_hal_beg = (int)plan_halo.size();
    ch_w = ch_wo = ch_wm_est = 0;
  };
  // distinct outside columns row i adds to the open chunk (commit: give them slots)
  auto new_halo_of = [&](int64_t i, int64_t r0, int64_t r1, int chunk_id, bool commit) {
    int add = 0;
    const int mark = commit ? chunk_id : -2 - chunk_id;  // dry-run marks never equal a chunk id
    for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
      const int j = cols[k];
      if (j == i || !(backward ? (j > i) : (j < i)) || (j >= r0 && j < r1)) continue;
      if (seen[j] == chunk_id || seen[j] == mark) continue;
      ++add;
      seen[j] = mark;
      if (commit) plan_halo.push_back(j);
    }
    if (!commit)  // undo the marks of the dry run
      for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
        const int j = cols[k];
        if (j >= 0 && j < n && seen[j] == mark) seen[j] = -1;
      }
    return add;
  };

  for (int64_t t = 0; t < nb; ++t) {
    const int64_t b = backward ? nb - 1 - t : t;
    const int64_t r0 = b * R, r1 = std::min(n, r0 + R);
    blk_chunk[t] = (int)plans.size();
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
        const RowInfo ri = rinfo[i];
        if (wave_layout<T>(1, 1, pad4(ri.np), ri.nx, ext_out[i]).total > ws.stage_static || ri.no > ws.stage_other) return;  // row too long for a stage
        const int nr1 = chunk_rows() + 1;
        const int nseg1 = (int)plan_segs.size() - ch_seg_beg + 1;
        const int w1 = std::max(ch_w, ri.np), wo1 = std::max(ch_wo, ri.no), wm1 = std::max(ch_wm_est, ext_out[i]);
        // ri.nx over-counts the new slots when the chunk already holds some of the columns: a
        // conservative test first, the exact count only when that one fails
        bool fits = chunk_rows() == 0;
        if (!fits && nr1 <= ws.stage_rows && (backward || (int64_t)wo1 * nr1 <= ws.stage_other)) {
          fits = wave_layout<T>(nr1, nseg1, pad4(w1), chunk_halo() + ri.nx, wm1).total <= ws.stage_static;
          if (!fits && ri.nx > 0) {
            const int add = new_halo_of(i, r0, r1, (int)plans.size(), false);
            fits = wave_layout<T>(nr1, nseg1, pad4(w1), chunk_halo() + add, wm1).total <= ws.stage_static;
          }
        }
        if (!fits) close_chunk(kblk++);
        new_halo_of(i, r0, r1, (int)plans.size(), true);
        plan_rows.push_back((int)i);
        ch_w = std::max(ch_w, ri.np);
        ch_wo = std::max(ch_wo, ri.no);
        ch_wm_est = std::max(ch_wm_est, ext_out[i]);
        seg_open = true;
      }
      if (seg_open) {  // end of a level: barrier point
        plan_segs.push_back(chunk_rows());
        seg_open = false;
      }
    }
    close_chunk(kblk++);
    if (plans.size() >= (size_t)1 << 30) return;
  }
  if (mb_total >= ((long long)1 << 32) - 16) return;  // mailbox slots are addressed with 32 bits

  lap("pass A");
  // ---- mail lists: which slots want the value of row j ----------------------------------------
  std::vector<int64_t> mail_ptr(n + 1, 0);
  for (const Plan& pl : plans)
    for (int q = pl.hal_beg; q < pl.hal_end; ++q) mail_ptr[plan_halo[q] + 1]++;
  for (int64_t j = 0; j < n; ++j) mail_ptr[j + 1] += mail_ptr[j];
  std::vector<uint32_t> mail_idx((size_t)mail_ptr[n]);
  {
    std::vector<int64_t> cur(mail_ptr.begin(), mail_ptr.end() - 1);
    for (const Plan& pl : plans)
      for (int q = pl.hal_beg; q < pl.hal_end; ++q) mail_idx[(size_t)cur[plan_halo[q]]++] = (uint32_t)(pl.mb_off + (q - pl.hal_beg));
  }

  lap("mail lists");
  // ---- pass B: emit the packed chunks.  Output extents first (cheap, sequential), then the
  //      chunks are filled by a few host threads: every chunk owns disjoint ranges of every array.
  const size_t nplans = plans.size();
  std::vector<size_t> stat_off(nplans + 1, 0);
  std::vector<int64_t> slot_off(nplans + 1, 0), auxo(nplans, 0);
  std::vector<int> wm_of(nplans, 0);
  for (size_t ci = 0; ci < nplans; ++ci) {
    const Plan& pl = plans[ci];
    const int nrows = pl.row_end - pl.row_beg, nseg = pl.seg_end - pl.seg_beg, nhalo = pl.hal_end - pl.hal_beg;
    int Wm = 0;
    for (int q = 0; q < nrows; ++q) {
      const int64_t i = plan_rows[pl.row_beg + q];
      Wm = std::max<int>(Wm, (int)(mail_ptr[i + 1] - mail_ptr[i]));
    }
    wm_of[ci] = Wm;
    const WaveLayout L = wave_layout<T>(nrows, nseg, pl.W, nhalo, Wm);
    if (L.total > ws.stage_static) return;  // cannot happen: Wm <= the estimate used in pass A
    stat_off[ci + 1] = stat_off[ci] + (size_t)L.total;
    slot_off[ci + 1] = align_slots(slot_off[ci] + nrows);
    if (backward) {
      auxo[ci] = slot_off[ci];  // one pre-folded value per row, same slots as rhs
    } else {
      auxo[ci] = aux_slots;
      aux_slots = align_slots(aux_slots + (int64_t)pl.Wo * nrows);
    }
  }
  stat.assign(stat_off[nplans], 0);
  chunks.assign(nplans, WaveChunk{});
  rowmap.assign((size_t)slot_off[nplans], -1);
  if (!backward) {
    aux_base.assign((size_t)slot_off[nplans], 0);
    aux_dims.assign(2 * (size_t)slot_off[nplans], 0);
  }
  auto emit_range = [&](size_t c_begin, size_t c_end) {
    std::vector<std::pair<int, int>> halo_slot;  // (column, slot) of the chunk, sorted by column
    for (size_t ci = c_begin; ci < c_end; ++ci) {
      const Plan& pl = plans[ci];
      const int nrows = pl.row_end - pl.row_beg, nseg = pl.seg_end - pl.seg_beg, nhalo = pl.hal_end - pl.hal_beg;
      const int W = pl.W, Wo = pl.Wo, Wm = wm_of[ci];
      const int64_t first = plan_rows[pl.row_beg];
      const int64_t r0 = (first / R) * R, r1 = std::min(n, r0 + R);
      const WaveLayout L = wave_layout<T>(nrows, nseg, W, nhalo, Wm);
      const size_t base = stat_off[ci];
      const uint32_t stage_base = (uint32_t)(ring_off + (size_t)(pl.kblk % ws.stages) * stage_bytes);
      WaveChunk d{};
      d.soff = (long long)base;
      d.sbytes = L.total;
      d.nrows = nrows;
      d.rhs_off = (long long)slot_off[ci];
      d.aux_off = (long long)auxo[ci];
      d.aux_cnt = backward ? nrows : Wo * nrows;
      halo_slot.clear();
      for (int h = 0; h < nhalo; ++h) halo_slot.emplace_back(plan_halo[pl.hal_beg + h], h);
      std::sort(halo_slot.begin(), halo_slot.end());
      std::vector<uint32_t> eoff((size_t)W * nrows, (uint32_t)WAVE_ZERO_OFF), mail((size_t)Wm * nrows, SPB_WAVE_NOMAIL);
      std::vector<T> ev((size_t)W * nrows, zero_of<T>()), dg(nrows), hs(nhalo);
      for (int h = 0; h < nhalo; ++h) {
        unsigned long long w[2] = {sentinel, sentinel};
        memcpy(&hs[h], w, sizeof(T));
      }
      for (int q = 0; q < nrows; ++q) {
        const int64_t i = plan_rows[pl.row_beg + q];
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
            const auto it = std::lower_bound(halo_slot.begin(), halo_slot.end(), std::make_pair(j, 0));
            off = stage_base + (uint32_t)L.hslot + (uint32_t)(it->second * sizeof(T));
          }
          eoff[(size_t)e * nrows + q] = off;
          ev[(size_t)e * nrows + q] = vals[k];
          ++e;
        }
        dg[q] = dv;
        int m = 0;
        for (int64_t k = mail_ptr[i]; k < mail_ptr[i + 1]; ++k) mail[(size_t)(m++) * nrows + q] = mail_idx[(size_t)k];
        const size_t slot = (size_t)slot_off[ci] + q;
        rowmap[slot] = (int)i;
        if (!backward) {
          aux_base[slot] = d.aux_off + q;
          aux_dims[2 * slot] = nrows;
          aux_dims[2 * slot + 1] = Wo;
        }
      }
      const int hdr[16] = {nrows, nseg, W, nhalo, Wo, Wm, (int)(unsigned)(pl.mb_off & 0xffffffffLL), (int)(pl.mb_off >> 32),
                           L.seg_end, L.rowid, L.diag, L.eoff, L.eval, L.mail, L.hslot, L.hcol};
      put_bytes(stat, base, hdr, 16);
      put_bytes(stat, base + L.seg_end, plan_segs.data() + pl.seg_beg, nseg);
      put_bytes(stat, base + L.rowid, plan_rows.data() + pl.row_beg, nrows);
      put_bytes(stat, base + L.diag, dg.data(), nrows);
      put_bytes(stat, base + L.eoff, eoff.data(), eoff.size());
      put_bytes(stat, base + L.eval, ev.data(), ev.size());
      put_bytes(stat, base + L.mail, mail.data(), mail.size());
      put_bytes(stat, base + L.hslot, hs.data(), hs.size());
      {  // where every landing slot's value is produced: block << 16 | row inside the block (cluster mode)
        std::vector<unsigned> hw(nhalo);
        for (int h = 0; h < nhalo; ++h) {
          const int64_t j = plan_halo[pl.hal_beg + h];
          hw[h] = (unsigned)((j / R) << 16) | (unsigned)(j % R);
        }
        put_bytes(stat, base + L.hcol, hw.data(), nhalo);
      }
      chunks[ci] = d;
    }
  };
  {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::max<size_t>(1, std::min<size_t>({(size_t)8, (size_t)hw / 2 + 1, nplans / 64 + 1}));
    std::vector<std::thread> pool;
    for (size_t k = 1; k < nt; ++k) pool.emplace_back(emit_range, nplans * k / nt, nplans * (k + 1) / nt);
    emit_range(0, nplans / nt);
    for (auto& th : pool) th.join();
  }
  lap("pass B");
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
  ws.mailbox_slots = mb_total;
  ws.mailbox.alloc(sizeof(T) * ((size_t)mb_total + 4));
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
  lap("upload");
  ws.ok = true;
}

template <typename T, typename IP>
static void wave_prep_launch(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, const T* sig) {
  Ctx* c = M->ctx;
  CsrMat<T>* A = M->A;
  const int64_t mb8 = (ws.mailbox_slots * (int64_t)sizeof(T) + 7) / 8;  // (the allocation carries 4 spare slots)
  const int64_t work = std::max(mb8, ws.rhs_slots);
  LaunchScope lsc(c, FAM_PRECOND);
  auto kern = ws.backward ? gs_wave_prep_kernel<T, IP, true> : gs_wave_prep_kernel<T, IP, false>;
  kern<<<vec_grid(c, work), kVecThreads, 0, c->stream>>>(mb8, reinterpret_cast<unsigned long long*>(ws.mailbox.p), ws.rhs_slots, bufptr<int>(ws.rowmap),
                                                          rhs, bufptr<T>(ws.rhsp), other, bufptr<IP>(A->indptr), bufptr<int>(A->cols),
                                                          bufptr<T>(A->vals), bufptr<T>(ws.aux), bufptr<long long>(ws.aux_base),
                                                          bufptr<int>(ws.aux_dims), bufptr<int>(ws.ticket), c->gate, c->gate_value, sig);
  check_launch("gs_wave_prep_kernel");
}

template <typename T>
void wave_sweep(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, T* out, T* sig) {
  Ctx* c = M->ctx;
  const int64_t n = M->A->n_local;
  if (n == 0) return;
  const int rhs_bytes = wave_a16((long long)sizeof(T) * ws.stage_rows);
  const int oth_bytes = wave_a16((long long)sizeof(T) * ws.stage_other);
  const int stage_bytes = (ws.stage_static + rhs_bytes + oth_bytes + 127) / 128 * 128;
  auto prep = [&]() {
    if (M->A->ip64)
      wave_prep_launch<T, int64_t>(M, ws, rhs, other, ws.backward ? sig : nullptr);
    else
      wave_prep_launch<T, int32_t>(M, ws, rhs, other, ws.backward ? sig : nullptr);
  };
  WaveArgs<T> a{};
  a.stat = bufptr<unsigned char>(ws.stat);
  a.chunks = bufptr<WaveChunk>(ws.chunks);
  a.blk_chunk = bufptr<int>(ws.blk_chunk);
  a.rhsp = bufptr<T>(ws.rhsp);
  a.aux = other ? bufptr<T>(ws.aux) : nullptr;
  a.out = out;
  a.mailbox = bufptr<T>(ws.mailbox);
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
  a.err = c->dev_err;
  a.sig = ws.backward ? nullptr : sig;
  {
    const char* se = getenv("SPB_GS_SPIN_NS");
    a.spin_ns = se && *se ? (unsigned)atoi(se) : 0u;
  }
  // One sweep = the pre-pass + the wavefront kernel.  want > 0: cluster launch (consecutive blocks = CTAs of one cluster,
  // hand-offs through distributed shared memory) with `want` CTAs per cluster (16 = non-portable size) or the next
  // smaller size the device can co-schedule with this footprint; 0: plain CTAs + mailbox.  Returns the size used.
  auto run = [&](int want) -> int {
    prep();
    LaunchScope lsc(c, FAM_PRECOND);
    int csize = ws.nblocks < 2 ? 0 : want;
    while (csize >= 2) {
      auto kc = ws.backward ? gs_wave_kernel<T, true, true> : gs_wave_kernel<T, false, true>;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(ceil_div((int64_t)ws.nblocks, (int64_t)csize) * csize));
      cfg.blockDim = dim3(WAVE_THREADS);
      cfg.dynamicSmemBytes = ws.smem_bytes;
      cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)csize;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int ncl = 0;
      const bool ok = cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.smem_bytes) == cudaSuccess &&
                      (csize <= 8 || cudaFuncSetAttribute(kc, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) &&
                      cudaOccupancyMaxActiveClusters(&ncl, kc, &cfg) == cudaSuccess && ncl >= 1;
      if (ok) {
        SPB_CUDA(cudaLaunchKernelEx(&cfg, kc, a));
        check_launch("gs_wave_kernel (cluster)");
        return csize;
      }
      cudaGetLastError();
      csize /= 2;
    }
    auto kern = ws.backward ? gs_wave_kernel<T, true, false> : gs_wave_kernel<T, false, false>;
    // per launch, not cached: the attribute is per device and a process may hold contexts on several
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws.smem_bytes));
    kern<<<ws.nblocks, WAVE_THREADS, ws.smem_bytes, c->stream>>>(a);
    check_launch("gs_wave_kernel");
    return 0;
  };
  if (ws.cluster < 0) {
    // First sweep of this schedule: which launch shape is fastest depends on how the blocks map onto the GPCs
    // (measured on B200, 7-point: 100^3, 100 blocks: clusters of 16 are 13 % faster; 128^3, 128 blocks: clusters of 16 / 8 are
    // 2-5 % slower -- not all of them fit at once -- and clusters of 4 are 10 % faster).
    // A sweep is a pure function of its inputs (the pre-pass resets mailbox and ticket), so it is simply run with each
    // candidate and timed -- same bits whatever is chosen.  SPB_GS_CLUSTER fixes the choice.
    const char* ce = getenv("SPB_GS_CLUSTER");
    if (ce && *ce) {
      ws.cluster = std::max(0, atoi(ce));
    } else if (c->gate || ws.nblocks < 4) {
      ws.cluster = 0;  // (inside a gated launch queue the kernels may be no-ops: nothing to time)
    } else {
      cudaEvent_t e0, e1;
      SPB_CUDA(cudaEventCreate(&e0));
      SPB_CUDA(cudaEventCreate(&e1));
      float best_ms = 0.f;
      int best = 0, last_used = -1;
      for (int cand : {16, 8, 4, 2, 0}) {
        const int used = run(cand);  // warm-up, and what the device really grants
        if (used == last_used) continue;
        last_used = used;
        SPB_CUDA(cudaEventRecord(e0, c->stream));
        run(used);
        run(used);
        SPB_CUDA(cudaEventRecord(e1, c->stream));
        SPB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SPB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (best_ms == 0.f || ms < best_ms) {
          best_ms = ms;
          best = used;
        }
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      ws.cluster = best;
    }
  }
  run(ws.cluster);
}

#define SPB_INST_WAVE(T)                                                                                              \
  template void wave_build<T>(CsrMat<T>*, const std::vector<int64_t>&, const std::vector<int>&, const std::vector<T>&, \
                              bool, WaveSched&);                                                                      \
  template void wave_sweep<T>(GsOp<T>*, WaveSched&, const T*, const T*, T*, T*);
SPB_INST_WAVE(double)
SPB_INST_WAVE(cplx)
SPB_INST_WAVE(float)
SPB_INST_WAVE(cplxf)

}  // namespace spb
