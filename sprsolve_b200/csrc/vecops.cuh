// vecops.cuh -- device vector algebra: the vecalg module (src/vecalg.rs:19-144, fallbacks
// :556-605) on device pointers, plus the pieces the fused solver kernels share.
#pragma once
#include "common.cuh"
#include "reduce.cuh"

namespace spb {

static const int kVecThreads = 256;

inline int vec_grid(const Ctx* c, int64_t n) {
  // persistent-style grid: a multiple of the SM count, every CTA grid-strides
  const int64_t want = ceil_div(n, kVecThreads);
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}
inline int64_t vec_max_grid(const Ctx* c) { return (int64_t)c->sm_count * 8; }
// Grid of a vector kernel that does not fit 8 CTAs per SM (the kernels that also carry double-double partial sums use
// 40-48 registers: 5-6 resident CTAs): exactly ONE wave of resident CTAs.  With the fixed 8 per SM such a kernel ran
// 1.6 waves, the second one 60 % full (bicg_k3: 5.2 TB/s where bicg_k1 / k2 reach 6.2-6.3, profiles/r02_bicg_k_ncu_full.json).
template <typename K>
inline int vec_grid_resident(const Ctx* c, int64_t n, K kernel) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kVecThreads, 0) != cudaSuccess || nb < 1) nb = 8;
  const int64_t cap = (int64_t)c->sm_count * std::min(nb, 8);
  return (int)std::max<int64_t>(1, std::min(ceil_div(n, (int64_t)kVecThreads), cap));
}

#define SPB_GRID_STRIDE(i, n)                                                        \
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n);          \
       i += (int64_t)gridDim.x * blockDim.x)

// Per-block partial sums of up to two quantities -> partials[2*block + slot].
template <typename T>
__device__ __forceinline__ void write_partials(T e0, T e1, T* partials) {
  __shared__ T scratch[32];
  e0 = block_sum(e0, scratch);
  e1 = block_sum(e1, scratch);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = e0;
    partials[2 * blockIdx.x + 1] = e1;
  }
}

// Fixed-order double-double sum of the block partials, rounded once, into red[0], red[1] (one tiny
// CTA).  allreduce: followed by the sum over ranks before the rounding (fused into the same kernel
// on the peer transport; all-gather of the unrounded pairs on the NCCL transport).
template <typename T>
void finalize_reduce(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red, bool allreduce);
template <typename T>
void finalize_partials(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red);
template <typename T>
void finalize_allreduce(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red);

// vecalg on device pointers ---------------------------------------------------------------
// kind: 0 = dot (no conjugate, vecalg.rs:557-561), 1 = conj_dot (:564-568), 2 = sum |x|^2 (:601-605)
template <typename T>
void vec_reduce(Ctx* c, int kind, int64_t n, const T* x, const T* y, Acc<T>* partials, scal2* red, bool allreduce = false);
template <typename T>
void vec_axpy(Ctx* c, int64_t n, T a, const T* x, T* y);
template <typename T>
void vec_axpby(Ctx* c, int64_t n, T a, const T* x, T b, T* y);
template <typename T>
void vec_scale(Ctx* c, int64_t n, T a, T* x);
template <typename T>
void vec_rscale(Ctx* c, int64_t n, real_t<T> a, T* x);
template <typename T>
void vec_conj(Ctx* c, int64_t n, const T* x, T* out);
template <typename T>
void vec_zero(Ctx* c, int64_t n, T* x);

}  // namespace spb
