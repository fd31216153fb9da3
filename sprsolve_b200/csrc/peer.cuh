// peer.cuh -- device side of the NVLink peer-memory data path (one process per GPU, CUDA IPC).
//
// Every rank owns small "windows" of device memory that all ranks of the node have mapped
// (dist.cu: window_create).  Collectives are then plain stores into the peers' windows over
// NVLink / NVSwitch followed by a release flag, and spins on flags in local memory:
//   * the Krylov-scalar all-reduce is fused into the kernel that finishes the local reduction
//     (one tiny CTA: sum block partials -> store the local sums into every peer's slot -> wait for
//     every peer's flag -> add the W contributions in rank order, so all ranks get identical bits);
//   * the halo exchange is a "put": the pack kernel writes this rank's boundary x entries
//     straight into the neighbours' halo buffers, and the SpMV kernel itself waits for the
//     neighbours' flags right before its first boundary tile (spmv.cu).
// Sequence numbers are monotone and every buffer is double-buffered by sequence parity.  For the
// scalar all-reduce a rank can only be one step ahead of a peer (it needs every peer's contribution
// to finish its own step), so parity never collides.  The halo put has no such back-pressure by
// itself -- coupling may be one-directional (a rank that only sends never waits for anybody) -- so
// every receiver ACKNOWLEDGES: at the start of its own put s it tells each peer that its exchanges
// < s are consumed (its SpMV s-1 is complete in stream order), and a put waits until all peers have
// acknowledged exchange s-2, the previous user of the parity buffer it is about to overwrite.
// The reference has no analogue (single process).
#pragma once
#include <stdint.h>

namespace spb {

static const int kMaxPeers = 16;

struct PeerPtrs {
  void* p[kMaxPeers];  // window base of every rank, mapped into this process (p[rank] is local)
  int world;
  int rank;
};

// Scalar all-reduce window (one per context).
struct ScalWin {
  double slots[2][kMaxPeers][8];  // one reduction point: 4 (hi, lo) pairs -- slot 0 (re, im), slot 1 (re, im)
  unsigned long long flags[2][kMaxPeers];
  unsigned long long seq;
  int error;  // set when a spin timed out (a peer died): surfaces as SPB_NCCL_ERROR on the host
};

// Head of a halo window (one per partitioned matrix); the payload (2 x n_halo T, double
// buffered) follows at kHaloHeadBytes.
static const size_t kHaloHeadBytes = 512;
struct HaloHead {
  unsigned long long flags[kMaxPeers];  // flags[q] = sequence number of the last put of rank q
  unsigned long long acks[kMaxPeers];   // acks[q] = rank q has consumed all exchanges <= this number
  unsigned long long seq;               // exchanges this rank has started
  unsigned int done;                    // CTAs of the current put that have finished
  int error;
  int npeers;                           // ranks this rank exchanges with (static after set-up)
  int peer_rank[kMaxPeers];
};
static_assert(sizeof(HaloHead) <= kHaloHeadBytes, "HaloHead must fit kHaloHeadBytes");

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag >= target.  Gives up after ~10 s (a peer that never arrives must not hang the
// GPU): returns false and the caller records the error.
__device__ __forceinline__ bool spin_until_ge(const unsigned long long* flag, unsigned long long target) {
  if (ld_acquire_sys(flag) >= target) return true;
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
    for (int i = 0; i < 64; ++i)
      if (ld_acquire_sys(flag) >= target) return true;
    if (global_timer_ns() - t0 > 10000000000ULL) return false;
  }
}

// Sum a reduction point over all ranks; called by ONE CTA with >= kMaxPeers threads, all threads.
// loc (shared memory) holds this rank's 4 double-double pairs (hi, lo) x {slot 0 re, im, slot 1 re,
// im}; on return loc[0..3] hold the four sums ROUNDED to double (identical bits on every rank, and
// -- because the pairs carry the sums exactly -- identical for every partition of the rows).
__device__ __forceinline__ void peer_allreduce_dd(double* loc, const PeerPtrs& pp) {
  ScalWin* me = static_cast<ScalWin*>(pp.p[pp.rank]);
  const unsigned long long seq = *((volatile unsigned long long*)&me->seq) + 1;
  const int par = (int)(seq & 1);
  const int t = threadIdx.x;
  __syncthreads();  // loc complete, everybody has read seq
  if (t < pp.world) {
    ScalWin* dst = static_cast<ScalWin*>(pp.p[t]);
    volatile double* s = dst->slots[par][pp.rank];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = loc[k];
    __threadfence_system();
    st_release_sys(&dst->flags[par][pp.rank], seq);
    if (!spin_until_ge(&me->flags[par][t], seq)) me->error = 1;
  }
  __syncthreads();
  if (t < 4) {  // pair t: double-double sum over the ranks in rank order, then ONE rounding
    double hi = 0.0, lo = 0.0;
    for (int q = 0; q < pp.world; ++q) {
      const volatile double* s = me->slots[par][q];
      const double h = s[2 * t], l = s[2 * t + 1];
      lo += l;
      const double sum = hi + h;
      const double bb = sum - hi;
      lo += (hi - (sum - bb)) + (h - bb);
      hi = sum;
    }
    __syncwarp(0xfu);
    loc[t] = hi + lo;
  }
  __syncthreads();
  if (t == 0) *((volatile unsigned long long*)&me->seq) = seq;
  __syncthreads();
}
#endif  // __CUDACC__

}  // namespace spb
