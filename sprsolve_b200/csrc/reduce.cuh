// reduce.cuh -- deterministic block reductions (no float atomics anywhere in the library).
//
// Every long sum of the reference (vecalg::conj_dot / norm2, src/vecalg.rs:563-568,601-605) is a
// sequential left fold.  On the GPU a sum is split into per-thread partial folds, a fixed shuffle
// tree per warp, a fixed-order fold over warps, one partial per block written to memory and a
// final fixed-order pass over the block partials.  Grid sizes depend only on the problem, so the
// result is bit-reproducible run to run; it differs from the reference only by summation order.
#pragma once
#include "scalar.cuh"

namespace spb {

__device__ __forceinline__ double shfl_down(double v, int d) {
  return __shfl_down_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ cplx shfl_down(cplx v, int d) {
  return cplx{__shfl_down_sync(0xffffffffu, v.re, d), __shfl_down_sync(0xffffffffu, v.im, d)};
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = add(v, shfl_down(v, d));
  return v;  // valid in lane 0
}

// Sum over the block; result valid in thread 0.  `scratch` holds >= 32 T.  Ends with a barrier so
// scratch can be reused immediately.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T tot = zero_of<T>();
  if (threadIdx.x == 0) {
    for (int w = 0; w < nwarps; ++w) tot = add(tot, scratch[w]);
  }
  __syncthreads();
  return tot;
}

// Fixed-order sum of `count` partials with stride `stride` starting at `p`, by one block.
template <typename T>
__device__ __forceinline__ T block_sum_partials(const T* p, int64_t count, int64_t stride,
                                                T* scratch) {
  T acc = zero_of<T>();
  for (int64_t i = threadIdx.x; i < count; i += blockDim.x) acc = add(acc, p[i * stride]);
  return block_sum(acc, scratch);
}

}  // namespace spb
