// scalar.cuh -- scalar arithmetic shared by every kernel of libsprsolve_b200.
//
// The element-wise arithmetic of the reference (cauchy::Scalar over f64 / Complex64 / f32 / Complex32 --
// the reference is generic over all four and dispatches s/d/c/z, src/mkl_mat.rs:68-71,198-201,
// src/vecalg.rs:192-195) is reproduced operation for operation so that a kernel differs from the
// reference's sequential code only in the order of long summations:
//   * no FMA contraction: the library is compiled with -fmad=false (Rust never fuses a*b+c);
//   * complex multiply/divide use the num_complex 0.3 formulas (4 mul + 2 add; naive division);
//   * mul_real / from_real / square / abs follow cauchy 0.3.
// Reference call sites: src/vecalg.rs:556-605, src/mat.rs:100-105, src/precond.rs:20-29,48-52.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define SPB_HD __host__ __device__ __forceinline__

namespace spb {

// Complex<f64>, interleaved (re, im), 16-byte aligned so one value is one 128-bit load.
struct __align__(16) cplx {
  double re, im;
};

// Complex<f32>, interleaved (re, im): one value is one 64-bit load.
struct __align__(8) cplxf {
  float re, im;
};

template <typename T>
struct ScalarTraits;
template <>
struct ScalarTraits<double> {
  static constexpr bool is_complex = false;
  static constexpr int dtype = 0;
  using real = double;
};
template <>
struct ScalarTraits<cplx> {
  static constexpr bool is_complex = true;
  static constexpr int dtype = 1;
  using real = double;
};
template <>
struct ScalarTraits<float> {
  static constexpr bool is_complex = false;
  static constexpr int dtype = 2;
  using real = float;
};
template <>
struct ScalarTraits<cplxf> {
  static constexpr bool is_complex = true;
  static constexpr int dtype = 3;
  using real = float;
};
// T::Real of cauchy::Scalar: every real-valued solver quantity (norms, Givens scalars, tolerances,
// epsilon) is computed in this type, so the f32 solvers round exactly where the reference's do.
template <typename T>
using real_t = typename ScalarTraits<T>::real;
template <typename T>
SPB_HD real_t<T> eps_of() {  // T::Real::epsilon()
  return sizeof(real_t<T>) == 4 ? (real_t<T>)1.1920928955078125e-07 : (real_t<T>)2.220446049250313e-16;
}
SPB_HD double sqrt_r(double v) { return sqrt(v); }
SPB_HD float sqrt_r(float v) { return sqrtf(v); }
SPB_HD double fabs_r(double v) { return fabs(v); }
SPB_HD float fabs_r(float v) { return fabsf(v); }

template <typename T>
SPB_HD T zero_of();
template <>
SPB_HD double zero_of<double>() {
  return 0.0;
}
template <>
SPB_HD cplx zero_of<cplx>() {
  return cplx{0.0, 0.0};
}
template <typename T>
SPB_HD T one_of();
template <>
SPB_HD double one_of<double>() {
  return 1.0;
}
template <>
SPB_HD cplx one_of<cplx>() {
  return cplx{1.0, 0.0};
}
template <>
SPB_HD float zero_of<float>() {
  return 0.0f;
}
template <>
SPB_HD cplxf zero_of<cplxf>() {
  return cplxf{0.0f, 0.0f};
}
template <>
SPB_HD float one_of<float>() {
  return 1.0f;
}
template <>
SPB_HD cplxf one_of<cplxf>() {
  return cplxf{1.0f, 0.0f};
}
template <typename T>
SPB_HD T from_real(real_t<T> r);
template <>
SPB_HD double from_real<double>(double r) {
  return r;
}
template <>
SPB_HD cplx from_real<cplx>(double r) {
  return cplx{r, 0.0};
}
template <>
SPB_HD float from_real<float>(float r) {
  return r;
}
template <>
SPB_HD cplxf from_real<cplxf>(float r) {
  return cplxf{r, 0.0f};
}

SPB_HD double add(double a, double b) { return a + b; }
SPB_HD double sub(double a, double b) { return a - b; }
SPB_HD double mul(double a, double b) { return a * b; }
SPB_HD double divi(double a, double b) { return a / b; }
SPB_HD double neg(double a) { return -a; }
SPB_HD double conj_of(double a) { return a; }
SPB_HD double mul_real(double a, double r) { return a * r; }
SPB_HD double square(double a) { return a * a; }
SPB_HD double abs_of(double a) { return fabs(a); }
SPB_HD double re_of(double a) { return a; }
SPB_HD double im_of(double) { return 0.0; }

SPB_HD cplx add(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
SPB_HD cplx sub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
SPB_HD cplx mul(cplx a, cplx b) {
  return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
SPB_HD cplx divi(cplx a, cplx b) {
  const double ns = b.re * b.re + b.im * b.im;
  const double re = a.re * b.re + a.im * b.im;
  const double im = a.im * b.re - a.re * b.im;
  return cplx{re / ns, im / ns};
}
SPB_HD cplx neg(cplx a) { return cplx{-a.re, -a.im}; }
SPB_HD cplx conj_of(cplx a) { return cplx{a.re, -a.im}; }
SPB_HD cplx mul_real(cplx a, double r) { return cplx{a.re * r, a.im * r}; }
SPB_HD double square(cplx a) { return a.re * a.re + a.im * a.im; }
SPB_HD double abs_of(cplx a) { return hypot(a.re, a.im); }
SPB_HD double re_of(cplx a) { return a.re; }
SPB_HD double im_of(cplx a) { return a.im; }

SPB_HD float add(float a, float b) { return a + b; }
SPB_HD float sub(float a, float b) { return a - b; }
SPB_HD float mul(float a, float b) { return a * b; }
SPB_HD float divi(float a, float b) { return a / b; }
SPB_HD float neg(float a) { return -a; }
SPB_HD float conj_of(float a) { return a; }
SPB_HD float mul_real(float a, float r) { return a * r; }
SPB_HD float square(float a) { return a * a; }
SPB_HD float abs_of(float a) { return fabsf(a); }
SPB_HD float re_of(float a) { return a; }
SPB_HD float im_of(float) { return 0.0f; }

SPB_HD cplxf add(cplxf a, cplxf b) { return cplxf{a.re + b.re, a.im + b.im}; }
SPB_HD cplxf sub(cplxf a, cplxf b) { return cplxf{a.re - b.re, a.im - b.im}; }
SPB_HD cplxf mul(cplxf a, cplxf b) {
  return cplxf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
SPB_HD cplxf divi(cplxf a, cplxf b) {
  const float ns = b.re * b.re + b.im * b.im;
  const float re = a.re * b.re + a.im * b.im;
  const float im = a.im * b.re - a.re * b.im;
  return cplxf{re / ns, im / ns};
}
SPB_HD cplxf neg(cplxf a) { return cplxf{-a.re, -a.im}; }
SPB_HD cplxf conj_of(cplxf a) { return cplxf{a.re, -a.im}; }
SPB_HD cplxf mul_real(cplxf a, float r) { return cplxf{a.re * r, a.im * r}; }
SPB_HD float square(cplxf a) { return a.re * a.re + a.im * a.im; }
SPB_HD float abs_of(cplxf a) { return hypotf(a.re, a.im); }
SPB_HD float re_of(cplxf a) { return a.re; }
SPB_HD float im_of(cplxf a) { return a.im; }

// T * V for the Jacobi preconditioner, V = T or V = real (DiagPrecond<Complex64,f64>).
SPB_HD double mul_diag(double a, double d) { return a * d; }
SPB_HD cplx mul_diag(cplx a, cplx d) { return mul(a, d); }
SPB_HD cplx mul_diag(cplx a, double d) { return mul_real(a, d); }
SPB_HD float mul_diag(float a, float d) { return a * d; }
SPB_HD cplxf mul_diag(cplxf a, cplxf d) { return mul(a, d); }
SPB_HD cplxf mul_diag(cplxf a, float d) { return mul_real(a, d); }

// Any scalar carried across the C ABI or kept in device-side solver state is a (re, im) pair.
struct __align__(16) scal2 {
  double re, im;
};
SPB_HD scal2 to_scal2(double a) { return scal2{a, 0.0}; }
SPB_HD scal2 to_scal2(cplx a) { return scal2{a.re, a.im}; }
SPB_HD scal2 to_scal2(float a) { return scal2{(double)a, 0.0}; }
SPB_HD scal2 to_scal2(cplxf a) { return scal2{(double)a.re, (double)a.im}; }
template <typename T>
SPB_HD T from_scal2(scal2 s);
template <>
SPB_HD double from_scal2<double>(scal2 s) {
  return s.re;
}
template <>
SPB_HD cplx from_scal2<cplx>(scal2 s) {
  return cplx{s.re, s.im};
}
// (the f32 types: the pair holds values already rounded to float, so the casts are exact)
template <>
SPB_HD float from_scal2<float>(scal2 s) {
  return (float)s.re;
}
template <>
SPB_HD cplxf from_scal2<cplxf>(scal2 s) {
  return cplxf{(float)s.re, (float)s.im};
}
// A reduction result is rounded ONCE to the precision of T::Real: double stays, float is the
// (correctly rounded) double rounded to float (the exact-sum checker of the tests rounds the same way).
template <typename T>
SPB_HD double round_to_real(double v) {
  return sizeof(real_t<T>) == 4 ? (double)(float)v : v;
}

}  // namespace spb
