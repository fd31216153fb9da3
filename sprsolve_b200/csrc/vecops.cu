// vecops.cu -- the vecalg module on the device (src/vecalg.rs): dot / conj_dot / norm2 /
// scale / rscale / conj / axpy / axpby, element-wise arithmetic identical to the fallbacks at
// src/vecalg.rs:556-605, reductions deterministic (reduce.cuh).  HBM-bound streaming kernels:
// coalesced grid-stride loops over a grid that is a multiple of the SM count.
#include "finalize.cuh"

namespace spb {

template <typename T>
void finalize_reduce(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red, bool allreduce) {
  finalize_reduce_tail<T>(c, partials, nblocks, red, allreduce, NoTail{});
}

template <typename T>
void finalize_partials(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red) {
  finalize_reduce<T>(c, partials, nblocks, red, false);
}

template <typename T>
void finalize_allreduce(Ctx* c, const Acc<T>* partials, int64_t nblocks, scal2* red) {
  finalize_reduce<T>(c, partials, nblocks, red, true);
}

template <typename T, int KIND>
__global__ void __launch_bounds__(kVecThreads) vec_reduce_k(int64_t n, const T* x, const T* y, Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>();
  SPB_GRID_STRIDE(i, n) {
    if (KIND == 0)
      acc_prod(e0, x[i], y[i]);           // dot, vecalg.rs:557-561
    else if (KIND == 1)
      acc_prod(e0, conj_of(x[i]), y[i]);  // conj_dot, :564-568
    else
      acc_sq(e0, x[i]);                   // norm2 squared, :601-605
  }
  write_partials(e0, zero_of<Acc<T>>(), partials);
}

template <typename T>
void vec_reduce(Ctx* c, int kind, int64_t n, const T* x, const T* y, Acc<T>* partials, scal2* red, bool allreduce) {
  // one wave of resident CTAs (the complex instances hold 40-44 registers: vecops.cuh, vec_grid_resident)
  const int grid = kind == 0 ? vec_grid_resident(c, n, vec_reduce_k<T, 0>) : kind == 1 ? vec_grid_resident(c, n, vec_reduce_k<T, 1>)
                                                                                        : vec_grid_resident(c, n, vec_reduce_k<T, 2>);
  {
    LaunchScope ls(c, FAM_VEC);
    if (kind == 0)
      vec_reduce_k<T, 0><<<grid, kVecThreads, 0, c->stream>>>(n, x, y, partials);
    else if (kind == 1)
      vec_reduce_k<T, 1><<<grid, kVecThreads, 0, c->stream>>>(n, x, y, partials);
    else
      vec_reduce_k<T, 2><<<grid, kVecThreads, 0, c->stream>>>(n, x, y, partials);
    check_launch("vec_reduce");
  }
  finalize_reduce<T>(c, partials, grid, red, allreduce);
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_axpy_k(int64_t n, T a, const T* x, T* y) {
  SPB_GRID_STRIDE(i, n) y[i] = add(y[i], mul(x[i], a));  // vecalg.rs:571-575
}
template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_axpby_k(int64_t n, T a, const T* x, T b, T* y) {
  SPB_GRID_STRIDE(i, n) y[i] = add(mul(x[i], a), mul(y[i], b));  // vecalg.rs:586-590
}
template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_scale_k(int64_t n, T a, T* x) {
  SPB_GRID_STRIDE(i, n) x[i] = mul(x[i], a);  // vecalg.rs:593-595
}
template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_rscale_k(int64_t n, real_t<T> a, T* x) {
  SPB_GRID_STRIDE(i, n) x[i] = mul_real(x[i], a);  // vecalg.rs:597-599
}
template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_conj_k(int64_t n, const T* x, T* out) {
  SPB_GRID_STRIDE(i, n) out[i] = conj_of(x[i]);  // vecalg.rs:578-583
}
template <typename T>
__global__ void __launch_bounds__(kVecThreads) vec_zero_k(int64_t n, T* x) {
  SPB_GRID_STRIDE(i, n) x[i] = zero_of<T>();
}

#define SPB_VEC_LAUNCH(kernel, ...)                                           \
  do {                                                                        \
    LaunchScope ls(c, FAM_VEC);                                               \
    kernel<<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(__VA_ARGS__);       \
    check_launch(#kernel);                                                    \
  } while (0)

template <typename T>
void vec_axpy(Ctx* c, int64_t n, T a, const T* x, T* y) {
  SPB_VEC_LAUNCH(vec_axpy_k<T>, n, a, x, y);
}
template <typename T>
void vec_axpby(Ctx* c, int64_t n, T a, const T* x, T b, T* y) {
  SPB_VEC_LAUNCH(vec_axpby_k<T>, n, a, x, b, y);
}
template <typename T>
void vec_scale(Ctx* c, int64_t n, T a, T* x) {
  SPB_VEC_LAUNCH(vec_scale_k<T>, n, a, x);
}
template <typename T>
void vec_rscale(Ctx* c, int64_t n, real_t<T> a, T* x) {
  SPB_VEC_LAUNCH(vec_rscale_k<T>, n, a, x);
}
template <typename T>
void vec_conj(Ctx* c, int64_t n, const T* x, T* out) {
  SPB_VEC_LAUNCH(vec_conj_k<T>, n, x, out);
}
template <typename T>
void vec_zero(Ctx* c, int64_t n, T* x) {
  SPB_VEC_LAUNCH(vec_zero_k<T>, n, x);
}

#define SPB_INST(T)                                                                          \
  template void finalize_reduce<T>(Ctx*, const Acc<T>*, int64_t, scal2*, bool);              \
  template void finalize_partials<T>(Ctx*, const Acc<T>*, int64_t, scal2*);                  \
  template void finalize_allreduce<T>(Ctx*, const Acc<T>*, int64_t, scal2*);                 \
  template void vec_reduce<T>(Ctx*, int, int64_t, const T*, const T*, Acc<T>*, scal2*, bool); \
  template void vec_axpy<T>(Ctx*, int64_t, T, const T*, T*);                                 \
  template void vec_axpby<T>(Ctx*, int64_t, T, const T*, T, T*);                             \
  template void vec_scale<T>(Ctx*, int64_t, T, T*);                                          \
  template void vec_rscale<T>(Ctx*, int64_t, real_t<T>, T*);                                 \
  template void vec_conj<T>(Ctx*, int64_t, const T*, T*);                                    \
  template void vec_zero<T>(Ctx*, int64_t, T*);
SPB_INST(double)
SPB_INST(cplx)
SPB_INST(float)
SPB_INST(cplxf)

}  // namespace spb
