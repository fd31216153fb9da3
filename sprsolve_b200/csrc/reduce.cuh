// reduce.cuh -- deterministic, order-independent block reductions (no float atomics anywhere).
//
// Every long sum of the reference (vecalg::conj_dot / norm2, src/vecalg.rs:563-568,601-605) is a
// sequential left fold.  On the GPU a sum is split into per-thread partial folds, a fixed shuffle
// tree per warp, a fixed-order fold over warps, one partial per block written to memory and a
// final fixed-order pass over the block partials (and, when partitioned, over the ranks).
//
// A plain double sum would make the RESULT depend on that split: on the launch plan of the SpMV
// (fused epilogue dots), on the grid size and on the number of GPUs -- and Jacobi-BiCGStab turns a
// 1e-16 difference of one dot product into a different iteration count (615..668 iterations on the
// 512^3 system, measured).  So the sums are carried in double-double (Acc<T>): every product is
// split exactly (p = a*b rounded, e = fma(a, b, -p)), partial sums are combined with error-free
// two-sums, and only the final value is rounded to double.  The rounded result is the correctly
// rounded exact sum except in astronomically rare near-tie cases, hence the same bits for every
// plan, grid and GPU count; it differs from the reference's sequential fold by that fold's own
// rounding error (1e-16 * sqrt(n) relative), i.e. exactly as a re-ordered sum would.
#pragma once
#include "scalar.cuh"

namespace spb {

__device__ __forceinline__ double shfl_down(double v, int d) {
  return __shfl_down_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ cplx shfl_down(cplx v, int d) {
  return cplx{__shfl_down_sync(0xffffffffu, v.re, d), __shfl_down_sync(0xffffffffu, v.im, d)};
}
__device__ __forceinline__ float shfl_down(float v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ cplxf shfl_down(cplxf v, int d) {
  return cplxf{__shfl_down_sync(0xffffffffu, v.re, d), __shfl_down_sync(0xffffffffu, v.im, d)};
}

// ---- double-double accumulators -----------------------------------------------------------------
// One real and one complex accumulator type, shared by the f64 and the f32 scalars: a product of two
// floats is exact in double, so the f32 types feed the same (hi, lo) pairs and are rounded to float
// at the very end (scalar.cuh: round_to_real).
struct __align__(16) AccR {
  double hi, lo;
};
struct __align__(16) AccC {
  double rh, rl, ih, il;
};
template <typename T>
struct AccOf {
  using type = AccR;
};
template <>
struct AccOf<cplx> {
  using type = AccC;
};
template <>
struct AccOf<cplxf> {
  using type = AccC;
};
template <typename T>
using Acc = typename AccOf<T>::type;
template <>
SPB_HD AccR zero_of<AccR>() {
  return AccR{0.0, 0.0};
}
template <>
SPB_HD AccC zero_of<AccC>() {
  return AccC{0.0, 0.0, 0.0, 0.0};
}
// (hi, lo) += x, x an exact double: error-free two-sum on hi, the error goes to lo
__device__ __forceinline__ void dd_add(double& hi, double& lo, double x) {
  const double s = hi + x;
  const double bb = s - hi;
  lo += (hi - (s - bb)) + (x - bb);
  hi = s;
}
// (hi, lo) += a * b exactly (the product error comes from one fma)
__device__ __forceinline__ void dd_add_prod(double& hi, double& lo, double a, double b) {
  const double p = a * b;
  lo += __fma_rn(a, b, -p);
  dd_add(hi, lo, p);
}
__device__ __forceinline__ void acc_prod(AccR& a, double x, double y) { dd_add_prod(a.hi, a.lo, x, y); }
__device__ __forceinline__ void acc_prod(AccC& a, cplx x, cplx y) {  // a += x * y
  dd_add_prod(a.rh, a.rl, x.re, y.re);
  dd_add_prod(a.rh, a.rl, -x.im, y.im);
  dd_add_prod(a.ih, a.il, x.re, y.im);
  dd_add_prod(a.ih, a.il, x.im, y.re);
}
__device__ __forceinline__ void acc_sq(AccR& a, double x) { dd_add_prod(a.hi, a.lo, x, x); }  // a += |x|^2
__device__ __forceinline__ void acc_sq(AccC& a, cplx x) {
  dd_add_prod(a.rh, a.rl, x.re, x.re);
  dd_add_prod(a.rh, a.rl, x.im, x.im);
}
// f32 operands: the product is exact in double (24 + 24 significand bits), no error term
__device__ __forceinline__ void acc_prod(AccR& a, float x, float y) { dd_add(a.hi, a.lo, (double)x * (double)y); }
__device__ __forceinline__ void acc_prod(AccC& a, cplxf x, cplxf y) {
  dd_add(a.rh, a.rl, (double)x.re * (double)y.re);
  dd_add(a.rh, a.rl, -(double)x.im * (double)y.im);
  dd_add(a.ih, a.il, (double)x.re * (double)y.im);
  dd_add(a.ih, a.il, (double)x.im * (double)y.re);
}
__device__ __forceinline__ void acc_sq(AccR& a, float x) { dd_add(a.hi, a.lo, (double)x * (double)x); }
__device__ __forceinline__ void acc_sq(AccC& a, cplxf x) {
  dd_add(a.rh, a.rl, (double)x.re * (double)x.re);
  dd_add(a.rh, a.rl, (double)x.im * (double)x.im);
}
__device__ __forceinline__ AccR add(AccR a, AccR b) {
  a.lo += b.lo;
  dd_add(a.hi, a.lo, b.hi);
  return a;
}
__device__ __forceinline__ AccC add(AccC a, AccC b) {
  a.rl += b.rl;
  dd_add(a.rh, a.rl, b.rh);
  a.il += b.il;
  dd_add(a.ih, a.il, b.ih);
  return a;
}
__device__ __forceinline__ AccR shfl_down(AccR v, int d) {
  return AccR{__shfl_down_sync(0xffffffffu, v.hi, d), __shfl_down_sync(0xffffffffu, v.lo, d)};
}
__device__ __forceinline__ AccC shfl_down(AccC v, int d) {
  return AccC{__shfl_down_sync(0xffffffffu, v.rh, d), __shfl_down_sync(0xffffffffu, v.rl, d),
              __shfl_down_sync(0xffffffffu, v.ih, d), __shfl_down_sync(0xffffffffu, v.il, d)};
}
// the four (hi, lo) pairs a reduction point carries: slot 0 (re, im), slot 1 (re, im)
__device__ __forceinline__ void acc_store(const AccR& a, double* dd4) {
  dd4[0] = a.hi;
  dd4[1] = a.lo;
  dd4[2] = 0.0;
  dd4[3] = 0.0;
}
__device__ __forceinline__ void acc_store(const AccC& a, double* dd4) {
  dd4[0] = a.rh;
  dd4[1] = a.rl;
  dd4[2] = a.ih;
  dd4[3] = a.il;
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
