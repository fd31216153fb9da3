// scalar.cuh -- scalar arithmetic shared by every kernel of libsprsolve_b200.
//
// The element-wise arithmetic of the reference (cauchy::Scalar over f64 / num_complex::Complex64)
// is reproduced operation for operation so that a kernel differs from the reference's sequential
// code only in the order of long summations:
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

template <typename T>
struct ScalarTraits;
template <>
struct ScalarTraits<double> {
  static constexpr bool is_complex = false;
  static constexpr int dtype = 0;
};
template <>
struct ScalarTraits<cplx> {
  static constexpr bool is_complex = true;
  static constexpr int dtype = 1;
};

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
template <typename T>
SPB_HD T from_real(double r);
template <>
SPB_HD double from_real<double>(double r) {
  return r;
}
template <>
SPB_HD cplx from_real<cplx>(double r) {
  return cplx{r, 0.0};
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

// T * V for the Jacobi preconditioner, V = T or V = real (DiagPrecond<Complex64,f64>).
SPB_HD double mul_diag(double a, double d) { return a * d; }
SPB_HD cplx mul_diag(cplx a, cplx d) { return mul(a, d); }
SPB_HD cplx mul_diag(cplx a, double d) { return mul_real(a, d); }

// Any scalar carried across the C ABI or kept in device-side solver state is a (re, im) pair.
struct __align__(16) scal2 {
  double re, im;
};
SPB_HD scal2 to_scal2(double a) { return scal2{a, 0.0}; }
SPB_HD scal2 to_scal2(cplx a) { return scal2{a.re, a.im}; }
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

#define SPB_EPS 2.220446049250313e-16 /* f64::EPSILON, T::Real::epsilon() in the reference */

}  // namespace spb
