// finalize.cuh -- the kernel that finishes a reduction point, with an optional fused TAIL.
//
// A Krylov iteration alternates "reduce" and "a few scalar operations on the result" (alpha =
// rho / <r0,v>, the Givens rotation ...).  The scalar step is a device functor executed by thread
// 0 of the finishing CTA right after the rounded sums are in memory, which saves one kernel launch
// per reduction point (the loops are latency-bound on small systems: ~3 us per launch).
#pragma once
#include "dist.cuh"
#include "vecops.cuh"

namespace spb {

struct NoTail {
  __device__ __forceinline__ void operator()() const {}
};

// Finishes a reduction point: fixed-order double-double sum of the block partials (reduce.cuh), the
// sum over the ranks, ONE rounding to double.  MODE 0: this rank only; MODE 1: the sum over ranks
// happens in the same CTA through the peers' scalar windows (peer.cuh); MODE 2 (NCCL transport):
// the unrounded pairs are written to dd_out, all-gathered, and finish_gathered_k sums and rounds.
template <typename T, int MODE, typename Tail>
__global__ void finalize_reduce_k(const Acc<T>* partials, int64_t nblocks, scal2* red, PeerPtrs pp, double* dd_out, Tail tail) {
  __shared__ Acc<T> scratch[32];
  __shared__ double loc[8];
  __shared__ double fin[4];
  const int t = threadIdx.x;
  for (int slot = 0; slot < 2; ++slot) {
    const Acc<T> s = block_sum_partials(partials + slot, nblocks, 2, scratch);
    if (t == 0) acc_store(s, loc + 4 * slot);
  }
  if (MODE == 1) {
    peer_allreduce_dd(loc, pp);
    if (t < 4) fin[t] = round_to_real<T>(loc[t]);
  } else {
    __syncthreads();
    if (MODE == 2) {
      if (t < 8) dd_out[t] = loc[t];
      return;
    }
    if (t < 4) fin[t] = round_to_real<T>(loc[2 * t] + loc[2 * t + 1]);
  }
  __syncthreads();
  if (t < 2) red[t] = scal2{fin[2 * t], fin[2 * t + 1]};
  __syncthreads();
  if (t == 0) tail();  // the scalar step that consumes red[] (fused: no extra launch)
}

// NCCL transport: gathered = [world][8]; rank-order double-double sum of every pair, one rounding.
template <typename T, typename Tail>
__global__ void finish_gathered_k(const double* gathered, int world, scal2* red, Tail tail) {
  __shared__ double fin[4];
  const int t = threadIdx.x;
  if (t < 4) {
    double hi = 0.0, lo = 0.0;
    for (int q = 0; q < world; ++q) {
      lo += gathered[8 * q + 2 * t + 1];
      dd_add(hi, lo, gathered[8 * q + 2 * t]);
    }
    fin[t] = round_to_real<T>(hi + lo);
  }
  __syncthreads();
  if (t < 2) red[t] = scal2{fin[2 * t], fin[2 * t + 1]};
  __syncthreads();
  if (t == 0) tail();
}

template <typename T, typename Tail>
void finalize_reduce_tail(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red, bool allreduce, Tail tail) {
  const bool dist = allreduce && c->dist && c->dist->world > 1;
  LaunchScope ls(c, FAM_SCALAR);
  if (!dist) {
    finalize_reduce_k<T, 0, Tail><<<1, 256, 0, c->stream>>>(partials, nblocks, red, PeerPtrs{}, nullptr, tail);
  } else if (peer_mode(c)) {
    finalize_reduce_k<T, 1, Tail><<<1, 256, 0, c->stream>>>(partials, nblocks, red, c->dist->scal->ptrs(), nullptr, tail);
  } else {
    Dist* d = c->dist;
    d->dd_buf.ensure(sizeof(double) * 8 * (size_t)(d->world + 1));
    double* send = bufptr<double>(d->dd_buf);
    double* recv = send + 8;
    finalize_reduce_k<T, 2, NoTail><<<1, 256, 0, c->stream>>>(partials, nblocks, red, PeerPtrs{}, send, NoTail{});
    SPB_NCCL(nccl().AllGather(send, recv, 8, ncclFloat64, d->comm, c->stream));
    finish_gathered_k<T, Tail><<<1, 32, 0, c->stream>>>(recv, d->world, red, tail);
  }
  check_launch("finalize_reduce");
}

}  // namespace spb
