// sprs_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).  See sprs_oracle.h.
//
// Every function cites the reference file:line it restates (paths relative to the reference
// crate root).  Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
// Arithmetic notes, all verified against the reference sources:
//   * Rust never contracts a*b+c into an FMA -> -ffp-contract=off here.
//   * num_complex 0.3 Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re); Div by complex:
//     (a.re*b.re + a.im*b.im)/|b|^2, (a.im*b.re - a.re*b.im)/|b|^2; Complex*real scales both parts.
//   * cauchy 0.3 Scalar: square() = re*re+im*im (real: x*x); abs() = hypot(re,im) (real: |x|);
//     mul_real scales both parts; from_real(x) = (x,0).
#include "sprs_oracle.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <omp.h>

#include <algorithm>
#include <vector>

namespace {

// 0 = serial everything (bit-defining oracle; = reference without `parallel`/`mkl`),
// 1 = OpenMP row-parallel SpMV only (= reference `parallel` feature, rayon; same numerics),
// 2 = OpenMP SpMV + OpenMP vector ops (stand-in for the `mkl` iomp build; summation order differs),
// 3 = "exact-dot" flavour: as 2, but every dot / conj_dot / norm2 (and the b-norm of GaussSeidel::solve)
//     is the EXACT sum of the exact products, rounded once (Kulisch-style superaccumulator below).
//     Element-wise work (SpMV folds, axpy, Jacobi, GS sweeps) is untouched.  This is the one member
//     of the family "reference algorithm + some summation order" that does not depend on the order,
//     so a second implementation with exactly rounded sums (the GPU library) must reproduce its
//     residual history bit for bit over every iteration.
int g_mode = 0;

struct cplx {
  double re, im;
};
// f32 / Complex32 (cauchy::Scalar is implemented for all four types; MKL dispatches s/d/c/z,
// src/mkl_mat.rs:68-71).  T::Real is f32 there: every real-valued quantity of the f32 solvers
// (norms, Givens scalars, tolerances, epsilon) is computed in float below.
struct cplxf {
  float re, im;
};
template <typename T>
struct RealOf {
  using type = double;
};
template <>
struct RealOf<float> {
  using type = float;
};
template <>
struct RealOf<cplxf> {
  using type = float;
};
template <typename T>
using real_t = typename RealOf<T>::type;
template <typename T>
inline real_t<T> eps_of() {  // T::Real::epsilon()
  return std::numeric_limits<real_t<T>>::epsilon();
}

// ---------- scalar traits -------------------------------------------------------------------
inline double zero_of(double) { return 0.0; }
inline cplx zero_of(cplx) { return cplx{0.0, 0.0}; }
inline double one_of(double) { return 1.0; }
inline cplx one_of(cplx) { return cplx{1.0, 0.0}; }
inline double from_real(double, double r) { return r; }
inline cplx from_real(cplx, double r) { return cplx{r, 0.0}; }

inline double add(double a, double b) { return a + b; }
inline double sub(double a, double b) { return a - b; }
inline double mul(double a, double b) { return a * b; }
inline double divi(double a, double b) { return a / b; }
inline double neg(double a) { return -a; }
inline double conj_of(double a) { return a; }
inline double mul_real(double a, double r) { return a * r; }
inline double square(double a) { return a * a; }
inline double abs_of(double a) { return std::fabs(a); }
inline double re_of(double a) { return a; }
inline double im_of(double) { return 0.0; }

inline cplx add(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
inline cplx sub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
inline cplx mul(cplx a, cplx b) {
  return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
inline cplx divi(cplx a, cplx b) {
  double ns = b.re * b.re + b.im * b.im;
  double re = a.re * b.re + a.im * b.im;
  double im = a.im * b.re - a.re * b.im;
  return cplx{re / ns, im / ns};
}
inline cplx neg(cplx a) { return cplx{-a.re, -a.im}; }
inline cplx conj_of(cplx a) { return cplx{a.re, -a.im}; }
inline cplx mul_real(cplx a, double r) { return cplx{a.re * r, a.im * r}; }
inline double square(cplx a) { return a.re * a.re + a.im * a.im; }
inline double abs_of(cplx a) { return std::hypot(a.re, a.im); }
inline double re_of(cplx a) { return a.re; }
inline double im_of(cplx a) { return a.im; }
// T * V with V real while T complex (DiagPrecond<Complex64,f64>, precond.rs:50)
inline cplx mul(cplx a, double r) { return cplx{a.re * r, a.im * r}; }


inline float zero_of(float) { return 0.0f; }
inline cplxf zero_of(cplxf) { return cplxf{0.0f, 0.0f}; }
inline float one_of(float) { return 1.0f; }
inline cplxf one_of(cplxf) { return cplxf{1.0f, 0.0f}; }
inline float from_real(float, float r) { return r; }
inline cplxf from_real(cplxf, float r) { return cplxf{r, 0.0f}; }
inline float add(float a, float b) { return a + b; }
inline float sub(float a, float b) { return a - b; }
inline float mul(float a, float b) { return a * b; }
inline float divi(float a, float b) { return a / b; }
inline float neg(float a) { return -a; }
inline float conj_of(float a) { return a; }
inline float mul_real(float a, float r) { return a * r; }
inline float square(float a) { return a * a; }
inline float abs_of(float a) { return std::fabs(a); }
inline float re_of(float a) { return a; }
inline float im_of(float) { return 0.0f; }
inline cplxf add(cplxf a, cplxf b) { return cplxf{a.re + b.re, a.im + b.im}; }
inline cplxf sub(cplxf a, cplxf b) { return cplxf{a.re - b.re, a.im - b.im}; }
inline cplxf mul(cplxf a, cplxf b) {
  return cplxf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
inline cplxf divi(cplxf a, cplxf b) {
  float ns = b.re * b.re + b.im * b.im;
  float re = a.re * b.re + a.im * b.im;
  float im = a.im * b.re - a.re * b.im;
  return cplxf{re / ns, im / ns};
}
inline cplxf neg(cplxf a) { return cplxf{-a.re, -a.im}; }
inline cplxf conj_of(cplxf a) { return cplxf{a.re, -a.im}; }
inline cplxf mul_real(cplxf a, float r) { return cplxf{a.re * r, a.im * r}; }
inline float square(cplxf a) { return a.re * a.re + a.im * a.im; }
inline float abs_of(cplxf a) { return std::hypot(a.re, a.im); }
inline float re_of(cplxf a) { return a.re; }
inline float im_of(cplxf a) { return a.im; }
inline cplxf mul(cplxf a, float r) { return cplxf{a.re * r, a.im * r}; }

// ---------- SpMV: src/mat.rs:68-129 ---------------------------------------------------------
template <typename T>
inline T row_fold(const int32_t* idx, const T* a, int64_t st, int64_t en, const T* x) {
  // mat.rs:100-105: fold(T::zero(), |acc,(lid,ldat)| acc + v_in[lid] * ldat)
  T acc = zero_of(T());
  for (int64_t k = st; k < en; ++k) acc = add(acc, mul(x[idx[k]], a[k]));
  return acc;
}

template <typename T>
void spmv_serial(int64_t n, const int64_t* indptr, const int32_t* idx, const T* a, const T* x,
                 T* y) {
  // mat.rs:71 zero fill, then mat.rs:114-128 serial windows(2) loop
  for (int64_t i = 0; i < n; ++i) y[i] = zero_of(T());
  for (int64_t i = 0; i < n; ++i) y[i] = row_fold(idx, a, indptr[i], indptr[i + 1], x);
}

template <typename T>
void spmv_par(int64_t n, const int64_t* indptr, const int32_t* idx, const T* a, const T* x, T* y) {
  // mat.rs:85-107: par_windows(2).with_min_len(128): static row chunks of >= 128 rows.
#pragma omp parallel for schedule(static, 128)
  for (int64_t i = 0; i < n; ++i) y[i] = row_fold(idx, a, indptr[i], indptr[i + 1], x);
}

template <typename T>
void spmv(int64_t n, const int64_t* indptr, const int32_t* idx, const T* a, const T* x, T* y) {
  if (g_mode >= 1)
    spmv_par(n, indptr, idx, a, x, y);
  else
    spmv_serial(n, indptr, idx, a, x, y);
}

// ---------- exact sums (mode 3) ---------------------------------------------------------------
// Fixed-point accumulator over the whole double range: limb i carries the bits of weight
// 2^(32 i - 1074).  A double m * 2^(e-1075) is split over three limbs and added without any
// rounding; limbs are int64, so 2^30 additions fit before carries must be propagated.
struct SuperAcc {
  static constexpr int N = 72;
  int64_t limb[N];
  int64_t count;
  bool special;  // a NaN / infinity went in
  SuperAcc() { clear(); }
  void clear() {
    std::memset(limb, 0, sizeof(limb));
    count = 0;
    special = false;
  }
  void normalize() {
    for (int i = 0; i + 1 < N; ++i) {
      const int64_t c = limb[i] >> 32;  // floor
      limb[i] -= c * ((int64_t)1 << 32);
      limb[i + 1] += c;
    }
    count = 0;
  }
  inline void add(double x) {
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff);
    uint64_t m = bits & (((uint64_t)1 << 52) - 1);
    if (e == 0) {
      if (m == 0) return;
      e = 1;
    } else if (e == 0x7ff) {
      special = true;
      return;
    } else {
      m |= (uint64_t)1 << 52;
    }
    const int pos = e - 1, i = pos >> 5, sh = pos & 31;
    const unsigned __int128 v = (unsigned __int128)m << sh;
    const int64_t a0 = (int64_t)(uint32_t)v, a1 = (int64_t)(uint32_t)(v >> 32), a2 = (int64_t)(uint64_t)(v >> 64);
    if (bits >> 63) {
      limb[i] -= a0;
      limb[i + 1] -= a1;
      limb[i + 2] -= a2;
    } else {
      limb[i] += a0;
      limb[i + 1] += a1;
      limb[i + 2] += a2;
    }
    if (++count >= ((int64_t)1 << 29)) normalize();
  }
  // += a * b exactly: p = fl(a b), e = a b - p (one fma; exact barring underflow)
  __attribute__((target("fma"))) inline void add_prod(double a, double b) {
    const double p = a * b;
    const double e = __builtin_fma(a, b, -p);
    add(p);
    add(e);
  }
  void merge(const SuperAcc& o) {
    normalize();
    SuperAcc t = o;
    t.normalize();
    for (int i = 0; i < N; ++i) limb[i] += t.limb[i];
    special = special || o.special;
    normalize();
  }
  // the exact value rounded to nearest-even double
  double round() {
    if (special) return std::numeric_limits<double>::quiet_NaN();
    normalize();
    bool negative = limb[N - 1] < 0;
    if (negative) {
      for (int i = 0; i < N; ++i) limb[i] = -limb[i];
      normalize();
    }
    int k = N - 1;
    while (k >= 0 && limb[k] == 0) --k;
    if (k < 0) return 0.0;
    auto L = [&](int i) -> unsigned __int128 { return i >= 0 ? (unsigned __int128)(uint64_t)limb[i] : 0; };
    const unsigned __int128 v = (L(k) << 64) | (L(k - 1) << 32) | L(k - 2);
    bool sticky = false;
    for (int i = k - 3; i >= 0 && !sticky; --i) sticky = limb[i] != 0;
    int hb = 95;
    while (!((v >> hb) & 1)) --hb;  // limb[k] != 0  =>  hb >= 64
    const int shift = hb - 52;
    uint64_t mant = (uint64_t)(v >> shift);
    const unsigned __int128 rem = v & (((unsigned __int128)1 << shift) - 1);
    const unsigned __int128 half = (unsigned __int128)1 << (shift - 1);
    if (rem > half || (rem == half && (sticky || (mant & 1)))) ++mant;
    const double r = std::ldexp((double)mant, shift + 32 * (k - 2) - 1074);
    return negative ? -r : r;
  }
};

// Exact sum over i of f(i) contributions, in parallel chunks (exactness makes the split irrelevant).
template <typename F>
double exact_sum(int64_t n, F contribute) {
  SuperAcc total;
#pragma omp parallel
  {
    SuperAcc local;
#pragma omp for schedule(static) nowait
    for (int64_t i = 0; i < n; ++i) contribute(local, i);
#pragma omp critical(spb_oracle_exact_merge)
    total.merge(local);
  }
  return total.round();
}
// products of floats are exact in double
inline double ex(float v) { return (double)v; }
inline double ex(double v) { return v; }
template <typename R>
inline void exact_prod(SuperAcc& a, R x, R y) {
  if (sizeof(R) == 4)
    a.add(ex(x) * ex(y));
  else
    a.add_prod(ex(x), ex(y));
}
// kind: 0 dot (no conjugate), 1 conj_dot.  Real types.
inline double exact_dot(int64_t n, const double* x, const double* y, int) {
  return exact_sum(n, [&](SuperAcc& a, int64_t i) { exact_prod(a, x[i], y[i]); });
}
inline float exact_dot(int64_t n, const float* x, const float* y, int) {
  return (float)exact_sum(n, [&](SuperAcc& a, int64_t i) { exact_prod(a, x[i], y[i]); });
}
template <typename C>
inline C exact_dot_c(int64_t n, const C* x, const C* y, int kind) {
  using R = decltype(x->re);
  const R sg = kind == 1 ? R(-1) : R(1);  // conj_dot: conj(x) . y
  const double re = exact_sum(n, [&](SuperAcc& a, int64_t i) {
    exact_prod(a, x[i].re, y[i].re);
    exact_prod(a, (R)(-sg * x[i].im), y[i].im);
  });
  const double im = exact_sum(n, [&](SuperAcc& a, int64_t i) {
    exact_prod(a, x[i].re, y[i].im);
    exact_prod(a, (R)(sg * x[i].im), y[i].re);
  });
  return C{(R)re, (R)im};
}
inline cplx exact_dot(int64_t n, const cplx* x, const cplx* y, int kind) { return exact_dot_c(n, x, y, kind); }
inline cplxf exact_dot(int64_t n, const cplxf* x, const cplxf* y, int kind) { return exact_dot_c(n, x, y, kind); }
inline double exact_sumsq(int64_t n, const double* x) {
  return exact_sum(n, [&](SuperAcc& a, int64_t i) { exact_prod(a, x[i], x[i]); });
}
inline float exact_sumsq(int64_t n, const float* x) {
  return (float)exact_sum(n, [&](SuperAcc& a, int64_t i) { exact_prod(a, x[i], x[i]); });
}
inline double exact_sumsq(int64_t n, const cplx* x) {
  return exact_sum(n, [&](SuperAcc& a, int64_t i) {
    exact_prod(a, x[i].re, x[i].re);
    exact_prod(a, x[i].im, x[i].im);
  });
}
inline float exact_sumsq(int64_t n, const cplxf* x) {
  return (float)exact_sum(n, [&](SuperAcc& a, int64_t i) {
    exact_prod(a, x[i].re, x[i].re);
    exact_prod(a, x[i].im, x[i].im);
  });
}

// ---------- vecalg fallbacks: src/vecalg.rs:556-605 -----------------------------------------
template <typename T>
T dot_fb(int64_t n, const T* x, const T* y) {  // vecalg.rs:557-561
  if (g_mode == 3) return exact_dot(n, x, y, 0);
  T acc = zero_of(T());
  for (int64_t i = 0; i < n; ++i) acc = add(acc, mul(x[i], y[i]));
  return acc;
}
inline double conj_dot_omp(int64_t n, const double* x, const double* y) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
inline cplx conj_dot_omp(int64_t n, const cplx* x, const cplx* y) {
  double sr = 0, si = 0;
#pragma omp parallel for reduction(+ : sr, si) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    cplx t = mul(conj_of(x[i]), y[i]);
    sr += t.re;
    si += t.im;
  }
  return cplx{sr, si};
}
inline float conj_dot_omp(int64_t n, const float* x, const float* y) {
  float s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
inline cplxf conj_dot_omp(int64_t n, const cplxf* x, const cplxf* y) {
  float sr = 0, si = 0;
#pragma omp parallel for reduction(+ : sr, si) schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    cplxf t = mul(conj_of(x[i]), y[i]);
    sr += t.re;
    si += t.im;
  }
  return cplxf{sr, si};
}
template <typename T>
T conj_dot(int64_t n, const T* x, const T* y) {  // vecalg.rs:564-568
  if (g_mode == 3) return exact_dot(n, x, y, 1);
  if (g_mode >= 2) return conj_dot_omp(n, x, y);
  T acc = zero_of(T());
  for (int64_t i = 0; i < n; ++i) acc = add(acc, mul(conj_of(x[i]), y[i]));
  return acc;
}
template <typename T>
real_t<T> norm2(int64_t n, const T* x) {  // vecalg.rs:601-605 (the sum and the sqrt are T::Real)
  real_t<T> acc = 0;
  if (g_mode == 3) return std::sqrt(exact_sumsq(n, x));
  if (g_mode >= 2) {
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < n; ++i) acc += square(x[i]);
  } else {
    for (int64_t i = 0; i < n; ++i) acc = acc + square(x[i]);
  }
  return std::sqrt(acc);
}
template <typename T>
void axpy(int64_t n, T a, const T* x, T* y) {  // vecalg.rs:571-575: *y += *x * a
#pragma omp parallel for schedule(static) if (g_mode >= 2)
  for (int64_t i = 0; i < n; ++i) y[i] = add(y[i], mul(x[i], a));
}
template <typename T>
void axpby(int64_t n, T a, const T* x, T b, T* y) {  // vecalg.rs:586-590: *y = *x*a + *y*b
#pragma omp parallel for schedule(static) if (g_mode >= 2)
  for (int64_t i = 0; i < n; ++i) y[i] = add(mul(x[i], a), mul(y[i], b));
}
template <typename T>
void scale(int64_t n, T a, T* x) {  // vecalg.rs:593-595: *v *= a
  for (int64_t i = 0; i < n; ++i) x[i] = mul(x[i], a);
}
template <typename T>
void rscale(int64_t n, real_t<T> a, T* x) {  // vecalg.rs:597-599: *v = v.mul_real(a)
#pragma omp parallel for schedule(static) if (g_mode >= 2)
  for (int64_t i = 0; i < n; ++i) x[i] = mul_real(x[i], a);
}
template <typename T>
void conj_vec(int64_t n, const T* x, T* out) {  // vecalg.rs:578-583
#pragma omp parallel for schedule(static) if (g_mode >= 2)
  for (int64_t i = 0; i < n; ++i) out[i] = conj_of(x[i]);
}
template <typename T>
void copy_vec(int64_t n, const T* src, T* dst) {  // ptr::copy_nonoverlapping
  std::memcpy(dst, src, sizeof(T) * (size_t)n);
}
template <typename T>
void zero_vec(int64_t n, T* x) {
  for (int64_t i = 0; i < n; ++i) x[i] = zero_of(T());
}

// ---------- operators -----------------------------------------------------------------------
template <typename T>
struct Csr {
  int64_t n;
  const int64_t* indptr;
  const int32_t* idx;
  const T* a;
  void mul_vec(const T* x, T* y) const { spmv(n, indptr, idx, a, x, y); }
  // mat.rs:145-152: SpMV then conj_dot(v_in, v_out)
  T mul_vec_dot(const T* x, T* y) const {
    mul_vec(x, y);
    return conj_dot(n, x, y);
  }
};

// One gauss_seidel.rs:111-125 sweep body for a single row, full row, diagonal skipped.
template <typename T>
inline T gs_row(const Csr<T>& A, int64_t row, T rhs_v, T diag, const T* x) {
  T sigma = zero_of(T());
  for (int64_t k = A.indptr[row]; k < A.indptr[row + 1]; ++k) {
    int64_t col = A.idx[k];
    if (col != row) sigma = add(sigma, mul(A.a[k], x[col]));  // :116 sigma += val * x[col]
  }
  return divi(sub(rhs_v, sigma), diag);  // :123
}

// Find the diagonal entry of each row the way gauss_seidel.rs:62-78 does (last match wins
// is impossible for sorted unique columns; first sweep caches `diag`).  Returns -1 or the
// first offending row.
template <typename T>
int64_t gs_diagonals(const Csr<T>& A, T* diag) {
  for (int64_t row = 0; row < A.n; ++row) {
    bool found = false;
    T d = zero_of(T());
    for (int64_t k = A.indptr[row]; k < A.indptr[row + 1]; ++k)
      if (A.idx[k] == row) {
        d = A.a[k];
        found = true;
      }
    if (!found) return row;               // :72-74
    if (square(d) < eps_of<T>()) return row;      // :76-78
    diag[row] = d;
  }
  return -1;
}

// Level schedule of a triangular dependency pattern (modes >= 1 only: lets the timed CPU arm and
// the exact-dot flavour run the Gauss-Seidel sweeps on all cores).  Rows of one level do not
// depend on each other; every row still folds its own sigma in CSR order from the same operand
// values as the sequential sweep, so the result is bit-identical to it (tests/test_oracle_exact.py).
struct Levels {
  std::vector<int64_t> ptr;
  std::vector<int32_t> rows;
  template <typename T>
  void build(const Csr<T>& A, bool lower) {
    const int64_t n = A.n;
    std::vector<int32_t> lev(n, 0);
    int32_t maxl = -1;
    auto visit = [&](int64_t i) {
      int32_t l = 0;
      for (int64_t k = A.indptr[i]; k < A.indptr[i + 1]; ++k) {
        const int64_t j = A.idx[k];
        if (lower ? (j < i) : (j > i)) l = std::max(l, lev[j] + 1);
      }
      lev[i] = l;
      maxl = std::max(maxl, l);
    };
    if (lower)
      for (int64_t i = 0; i < n; ++i) visit(i);
    else
      for (int64_t i = n - 1; i >= 0; --i) visit(i);
    ptr.assign(maxl + 2, 0);
    rows.resize(n);
    for (int64_t i = 0; i < n; ++i) ptr[lev[i] + 1]++;
    for (int32_t l = 0; l <= maxl; ++l) ptr[l + 1] += ptr[l];
    std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[cur[lev[i]]++] = (int32_t)i;
  }
};

// gauss_seidel.rs:111-125 row body with the operands of the two triangles taken from different
// vectors: columns < row from `lo`, columns > row from `hi` (null: the entries multiply the zeros
// of a sweep that starts from x = 0 and are skipped).  Same fold order as gs_row.
template <typename T>
inline T gs_row_split(const Csr<T>& A, int64_t row, T rhs_v, T diag, const T* lo, const T* hi) {
  T sigma = zero_of(T());
  for (int64_t k = A.indptr[row]; k < A.indptr[row + 1]; ++k) {
    const int64_t col = A.idx[k];
    if (col < row) {
      if (lo) sigma = add(sigma, mul(A.a[k], lo[col]));
    } else if (col > row) {
      if (hi) sigma = add(sigma, mul(A.a[k], hi[col]));
    }
  }
  return divi(sub(rhs_v, sigma), diag);
}

template <typename T>
struct Precond {
  int kind;
  int64_t n;
  const Csr<T>* A;     // GS kinds
  T* dinv_t;           // ORC_PC_DIAG: reciprocal (T)
  real_t<T>* dinv_r;   // ORC_PC_DIAG_REAL: reciprocal (real)
  T* gs_diag;          // GS kinds: cached diagonal
  T* gs_tmp;           // GS_SYM: forward result
  const Levels* lev_f; // modes >= 1: level schedules of the lower / upper pattern
  const Levels* lev_b;
  real_t<T> omega;     // SOR / SSOR kinds: relaxation factor
  // Relaxed update (SOR): x_i <- (1 - omega) x_i + omega * g_i, g_i the Gauss-Seidel value of
  // gauss_seidel.rs:123.  The reference has no relaxation; this fixes the operation order for it:
  // two real-by-scalar products, one add.  omega == 1 never takes this path (plain sweep, same bits as before).
  inline T relax(T xold, T g) const {
    const real_t<T> one_m = real_t<T>(1) - omega;
    return add(mul_real(xold, one_m), mul_real(g, omega));
  }
  // xold: the value x_i is relaxed against (null: 0, a sweep from zero; unused when omega == 1)
  void level_sweep(const Levels& L, const T* rhs, const T* lo, const T* hi, T* out, const T* xold = nullptr,
                   bool relaxed = false) const {
    const int64_t nl = (int64_t)L.ptr.size() - 1;
    for (int64_t l = 0; l < nl; ++l) {
      const int64_t b = L.ptr[l], e = L.ptr[l + 1];
#pragma omp parallel for schedule(static) if (e - b > 2048)
      for (int64_t i = b; i < e; ++i) {
        const int64_t r = L.rows[i];
        const T g = gs_row_split(*A, r, rhs[r], gs_diag[r], lo, hi);
        out[r] = relaxed ? relax(xold ? xold[r] : zero_of(T()), g) : g;
      }
    }
  }
  // precond.rs:48-52: *r = (*v) * (*s)
  void apply(const T* in, T* out) const {
    switch (kind) {
      case ORC_PC_DIAG:
        for (int64_t i = 0; i < n; ++i) out[i] = mul(in[i], dinv_t[i]);
        break;
      case ORC_PC_DIAG_REAL:
        for (int64_t i = 0; i < n; ++i) out[i] = mul_real(in[i], dinv_r[i]);
        break;
      case ORC_PC_GS_FWD:
        if (lev_f) {
          level_sweep(*lev_f, in, out, nullptr, out);
          break;
        }
        zero_vec(n, out);
        for (int64_t r = 0; r < n; ++r) out[r] = gs_row(*A, r, in[r], gs_diag[r], out);
        break;
      case ORC_PC_GS_SYM:
        if (lev_f && lev_b) {
          level_sweep(*lev_f, in, gs_tmp, nullptr, gs_tmp);
          level_sweep(*lev_b, in, gs_tmp, out, out);
          break;
        }
        // forward sweep from zero, then the same row body over rows n-1..0, in place.
        zero_vec(n, out);
        for (int64_t r = 0; r < n; ++r) out[r] = gs_row(*A, r, in[r], gs_diag[r], out);
        for (int64_t r = n - 1; r >= 0; --r) out[r] = gs_row(*A, r, in[r], gs_diag[r], out);
        break;
      case ORC_PC_SOR_FWD:
        // relaxed forward sweep from zero: z_i = (1 - w) * 0 + w * g_i
        if (lev_f) {
          level_sweep(*lev_f, in, out, nullptr, out, nullptr, true);
          break;
        }
        zero_vec(n, out);
        for (int64_t r = 0; r < n; ++r) out[r] = relax(out[r], gs_row(*A, r, in[r], gs_diag[r], out));
        break;
      case ORC_PC_SSOR:
        // SSOR(w): relaxed forward sweep from zero, then the relaxed sweep over rows n-1..0 in place:
        // z = w (2 - w) (D + w U)^-1 D (D + w L)^-1 r, symmetric positive definite for 0 < w < 2, D > 0.
        if (lev_f && lev_b) {
          level_sweep(*lev_f, in, gs_tmp, nullptr, gs_tmp, nullptr, true);
          level_sweep(*lev_b, in, gs_tmp, out, out, gs_tmp, true);
          break;
        }
        zero_vec(n, out);
        for (int64_t r = 0; r < n; ++r) out[r] = relax(out[r], gs_row(*A, r, in[r], gs_diag[r], out));
        for (int64_t r = n - 1; r >= 0; --r) out[r] = relax(out[r], gs_row(*A, r, in[r], gs_diag[r], out));
        break;
      default:
        copy_vec(n, in, out);
    }
  }
};

template <typename T>
struct PrecondOwner {
  Precond<T> p;
  PrecondOwner(int kind, int64_t n, const Csr<T>* A, const void* data) {
    p.kind = kind;
    p.n = n;
    p.A = A;
    p.dinv_t = nullptr;
    p.dinv_r = nullptr;
    p.gs_diag = nullptr;
    p.gs_tmp = nullptr;
    p.lev_f = nullptr;
    p.lev_b = nullptr;
    p.omega = real_t<T>(1);
    ok = true;
    bad_row = -1;
    if (kind == ORC_PC_SOR_FWD || kind == ORC_PC_SSOR) {
      p.omega = *reinterpret_cast<const real_t<T>*>(data);
      if (p.omega == real_t<T>(1)) p.kind = kind = (kind == ORC_PC_SSOR ? ORC_PC_GS_SYM : ORC_PC_GS_FWD);
    }
    if (kind == ORC_PC_DIAG) {
      // precond.rs:20-24: diag_inv.push(V::one() / *v)
      p.dinv_t = new T[n];
      const T* d = reinterpret_cast<const T*>(data);
      for (int64_t i = 0; i < n; ++i) p.dinv_t[i] = divi(one_of(T()), d[i]);
    } else if (kind == ORC_PC_DIAG_REAL) {
      p.dinv_r = new real_t<T>[n];
      const real_t<T>* d = reinterpret_cast<const real_t<T>*>(data);
      for (int64_t i = 0; i < n; ++i) p.dinv_r[i] = real_t<T>(1) / d[i];
    } else if (kind == ORC_PC_GS_FWD || kind == ORC_PC_GS_SYM || kind == ORC_PC_SOR_FWD || kind == ORC_PC_SSOR) {
      p.gs_diag = new T[n];
      bad_row = gs_diagonals(*A, p.gs_diag);
      ok = bad_row < 0;
      if (ok && g_mode >= 1) {
        lev_f.build(*A, true);
        p.lev_f = &lev_f;
        if (kind == ORC_PC_GS_SYM || kind == ORC_PC_SSOR) {
          lev_b.build(*A, false);
          p.lev_b = &lev_b;
          p.gs_tmp = new T[n];
        }
      }
    }
  }
  Levels lev_f, lev_b;
  ~PrecondOwner() {
    delete[] p.dinv_t;
    delete[] p.dinv_r;
    delete[] p.gs_diag;
    delete[] p.gs_tmp;
  }
  bool ok;
  int64_t bad_row;
};

struct Hist {
  double* buf;
  int64_t cap;
  int64_t len;
  void put(int64_t k, double v) {
    if (buf && k < cap) buf[k] = v;
    if (k + 1 > len) len = k + 1;
  }
};

// ---------- BiCGStab: src/bicg_stab.rs:35-200 (solve), :204-366 (precond_solve) -------------
// One routine: `M == nullptr` follows `solve` (y aliases p, z aliases r; slots per :65-69),
// otherwise `precond_solve` (slots per :235-241).  Operation order is the reference's.
template <typename T>
int bicgstab(int64_t size, int64_t n_rhs, int64_t n_x, const Csr<T>& A, const Precond<T>* M,
             const T* rhs, T* x, int64_t max_iter, double tol, T* ws, int64_t* iters,
             double* resid, Hist& h) {
  using R = real_t<T>;
  const R EPS_T = eps_of<T>();
  const R tol_r = (R)tol;
  const int64_t n = n_rhs;
  if (n != size) return ORC_INCOMPATIBLE_FORMAT;  // :44 / :214
  if (n != n_x) return ORC_INCOMPATIBLE_FORMAT;   // :49 / :219
  const R rhs_norm = norm2(n, rhs);          // :55 / :225
  if (rhs_norm <= EPS_T) {                          // :56-60
    zero_vec(n, x);
    *iters = 0;
    *resid = rhs_norm;
    return ORC_OK;
  }
  const R tol2 = tol_r * rhs_norm;  // :61
  T *r, *r0, *y, *p, *v, *t, *z;
  if (M) {  // :235-241
    r = ws, r0 = ws + n, y = ws + 2 * n, p = ws + 3 * n, v = ws + 4 * n, t = ws + 5 * n,
    z = ws + 6 * n;
  } else {  // :65-69
    r = ws, r0 = ws + n, y = ws + 2 * n, v = ws + 3 * n, t = ws + 4 * n;
    p = y;
    z = r;
  }
  A.mul_vec(x, r);                       // :73 / :244
  axpy(n, neg(one_of(T())), rhs, r);     // :75 r = A*x - rhs
  copy_vec(n, r, r0);                    // :78
  const R r0_norm = norm2(n, r0);   // :80
  h.put(0, r0_norm / rhs_norm);
  if (r0_norm <= tol2) {                 // :81-83
    *iters = 0;
    *resid = r0_norm / rhs_norm;
    return ORC_OK;
  }
  R r0_norm_tol = r0_norm * EPS_T;    // :84-85
  r0_norm_tol = r0_norm_tol * r0_norm_tol;

  T rho = from_real(T(), r0_norm * r0_norm);  // :88
  if (M) {
    copy_vec(n, r, p);   // :261
    M->apply(p, y);      // :262
  } else {
    copy_vec(n, r, y);   // :91 (y is p)
  }
  A.mul_vec(y, v);                                   // :93 / :263
  T alpha = divi(rho, conj_dot(n, r0, v));           // :96 / :266
  axpy(n, neg(alpha), v, r);                         // :100 / :269
  if (M) M->apply(r, z);                             // :273
  A.mul_vec(z, t);                                   // :104 / :275
  T tmp = conj_dot(n, t, t);                         // :107 / :278
  T w = re_of(tmp) > 0 ? divi(conj_dot(n, t, r), tmp) : zero_of(T());  // :108-113
  axpy(n, neg(alpha), y, x);                         // :115 / :288
  axpy(n, neg(w), z, x);                             // :117 / :290
  axpy(n, neg(w), t, r);                             // :120 / :293

  for (int64_t its = 1; its < max_iter; ++its) {     // :122 / :295
    const R r_norm = norm2(n, r);               // :123
    h.put(its, r_norm / rhs_norm);
    if (r_norm <= tol2) {                            // :124-126
      *iters = its;
      *resid = r_norm / rhs_norm;
      return ORC_OK;
    }
    const T rho_old = rho;                           // :127
    rho = conj_dot(n, r0, r);                        // :128
    if (abs_of(rho) < r0_norm_tol) {                 // :131-145 restart
      A.mul_vec(x, r);
      axpy(n, neg(one_of(T())), rhs, r);
      copy_vec(n, r, r0);
      const R rn = norm2(n, r);
      rho = from_real(T(), rn * rn);
      r0_norm_tol = re_of(rho) * EPS_T * EPS_T;
    }
    const T beta = mul(divi(rho, rho_old), divi(alpha, w));  // :146 / :319
    axpby(n, mul(neg(beta), w), v, beta, p);                 // :155 / :324
    axpy(n, one_of(T()), r, p);                              // :156 / :325
    if (M) M->apply(p, y);                                   // :328
    A.mul_vec(y, v);                                         // :160 / :329
    tmp = conj_dot(n, r0, v);                                // :163 / :332
    if (abs_of(tmp) <= 0) {                                // :164-167
      *iters = its;
      return ORC_BREAKDOWN;
    }
    alpha = divi(rho, tmp);                                  // :169
    axpy(n, neg(alpha), v, r);                               // :172
    if (M) M->apply(r, z);                                   // :343
    A.mul_vec(z, t);                                         // :175 / :344
    tmp = conj_dot(n, t, t);                                 // :178
    w = re_of(tmp) > 0 ? divi(conj_dot(n, t, r), tmp) : zero_of(T());  // :179-186
    axpy(n, neg(alpha), y, x);                               // :188 / :355
    axpy(n, neg(w), z, x);                                   // :191 / :357
    axpy(n, neg(w), t, r);                                   // :196 / :362
  }
  *iters = max_iter;
  return ORC_INSUFFICIENT_ITER;  // :199 / :365
}

// ---------- MINRES: src/minres.rs:31-172 (solve), :178-341 (precond_solve) ------------------
// CSMINRES: src/cs_minres.rs:29-158, selected with `cs == true` (no preconditioner there).
template <typename T>
int minres(int64_t size, int64_t n_rhs, int64_t n_x, const Csr<T>& A, const Precond<T>* M,
           bool cs, const T* rhs, T* x, int64_t max_iter, double tol, T* ws, int64_t* iters,
           double* resid, Hist& h) {
  using R = real_t<T>;
  const R EPS_T = eps_of<T>();
  const R tol_r = (R)tol;
  const int64_t n = n_rhs;
  if (n != size) return ORC_INCOMPATIBLE_FORMAT;
  if (n != n_x) return ORC_INCOMPATIBLE_FORMAT;
  const R rhs_norm = norm2(n, rhs);  // :51
  if (rhs_norm <= EPS_T) {
    zero_vec(n, x);
    *iters = 0;
    *resid = rhs_norm;
    return ORC_OK;
  }
  const R threshold = tol_r * rhs_norm;  // :57
  T c = one_of(T()), c_old = one_of(T());   // :60-64
  R s = 0.0, s_old = 0;
  T eta = one_of(T());
  T* v_old = ws;            // :68-73 / :216-223 / cs_minres.rs:66-72
  T* v_new = ws + n;
  T* v = ws + 2 * n;
  T* p_old = ws + 3 * n;
  T* p_oold = ws + 4 * n;
  T* p = ws + 5 * n;
  T* w = ws + 6 * n;        // precond only
  T* w_new = ws + 7 * n;    // precond only
  T* tvec = ws + 6 * n;     // cs only (cs_minres.rs:72)

  copy_vec(n, rhs, v_new);                   // :77
  A.mul_vec(x, v_old);                       // :78
  axpy(n, neg(one_of(T())), v_old, v_new);   // :80 v_new = rhs - A*x
  R res_norm = norm2(n, v_new);         // :81 / :231
  R beta_new, beta_one;
  if (M) {
    M->apply(v_new, w_new);                  // :233
    T b2 = conj_dot(n, v_new, w_new);        // :235
    if (re_of(b2) < EPS_T || im_of(b2) > EPS_T * re_of(b2)) return ORC_INVALID_PRECOND;  // :236-244
    beta_new = std::sqrt(re_of(b2));         // :245
    beta_one = beta_new;
    const R ts = R(1) / beta_new;        // :248-250
    rscale(n, ts, v_new);
    rscale(n, ts, w_new);
  } else {
    beta_new = res_norm;                     // :82-84
    beta_one = beta_new;
    rscale(n, R(1) / beta_new, v_new);
  }
  zero_vec(n, v);      // :86-88
  zero_vec(n, p_old);
  zero_vec(n, p);

  for (int64_t its = 0; its < max_iter; ++its) {  // :90
    const R beta = beta_new;
    T* v_t = v_old;  // :92-96 pointer rotation
    v_old = v;
    v = v_new;
    v_new = v_t;
    if (M) {  // :259-265
      T* w_t = w;
      w = w_new;
      w_new = w_t;
    }
    T alpha;
    const T* q;  // the vector p is seeded from
    if (cs) {
      conj_vec(n, v, tvec);            // cs_minres.rs:99
      A.mul_vec(tvec, v_new);          // :101
      alpha = conj_dot(n, v, v_new);   // :103
      q = tvec;
    } else if (M) {
      alpha = A.mul_vec_dot(w, v_new);  // minres.rs:271
      q = w;
    } else {
      alpha = A.mul_vec_dot(v, v_new);  // minres.rs:116
      q = v;
    }
    axpy(n, from_real(T(), -beta), v_old, v_new);  // :117 / :272 / cs:104
    axpy(n, neg(alpha), v, v_new);                 // :118 / :273 / cs:105
    if (M) {
      M->apply(v_new, w_new);                      // :276
      T b2 = conj_dot(n, v_new, w_new);            // :278
      if (re_of(b2) < EPS_T || im_of(b2) > EPS_T * re_of(b2)) {  // :279-287
        *iters = its;
        return ORC_INVALID_PRECOND;
      }
      beta_new = std::sqrt(re_of(b2));             // :288
      const R ts = R(1) / beta_new;            // :289-291
      rscale(n, ts, v_new);
      rscale(n, ts, w_new);
    } else {
      beta_new = norm2(n, v_new);                  // :120 / cs:106
      rscale(n, R(1) / beta_new, v_new);            // :121 / cs:107
    }
    // Givens rotation: minres.rs:132-148 ; cs_minres.rs:119-134 (conjugations differ)
    const R r3 = s_old * beta;
    const T tr = cs ? mul_real(conj_of(c_old), beta) : mul_real(c_old, beta);
    const T r2 = add(mul_real(alpha, s), mul(c, tr));
    const T r1_hat = cs ? sub(mul(conj_of(c), alpha), mul_real(tr, s))
                        : sub(mul(c, alpha), mul_real(tr, s));
    const R r1_inv = R(1) / std::sqrt(square(r1_hat) + beta_new * beta_new);
    c_old = c;
    s_old = s;
    c = cs ? mul_real(conj_of(r1_hat), r1_inv) : mul_real(r1_hat, r1_inv);
    s = beta_new * r1_inv;
    // solution update: :151-162 / cs:137-148
    T* p_t = p_oold;
    p_oold = p_old;
    p_old = p;
    p = p_t;
    copy_vec(n, q, p);                               // :156 / :325 / cs:142
    axpy(n, neg(r2), p_old, p);                      // :158
    axpy(n, from_real(T(), -r3), p_oold, p);         // :159
    rscale(n, r1_inv, p);                            // :160
    axpy(n, mul_real(mul(c, eta), beta_one), p, x);  // :162
    res_norm *= std::fabs(s);                        // :164
    h.put(its, res_norm / rhs_norm);
    if (res_norm < threshold) {                      // :165-167
      *iters = its;
      *resid = res_norm / rhs_norm;
      return ORC_OK;
    }
    eta = mul_real(eta, -s);                         // :168
  }
  *iters = max_iter;
  return ORC_INSUFFICIENT_ITER;
}

// ---------- GaussSeidel::solve: src/gauss_seidel.rs:33-140 ----------------------------------
template <typename T>
int gauss_seidel(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
                 const Csr<T>& A, const T* rhs, T* x, int64_t max_iter, double eps, T* ws,
                 int64_t* iters, double* resid, Hist& h, double omega_d = 1.0) {
  using R = real_t<T>;
  const R EPS_T = eps_of<T>();
  // omega != 1: successive over-relaxation, x_i <- (1 - w) x_i + w g_i (not in the reference; same
  // update as Precond::relax); omega == 1 is the reference's loop, untouched.
  const R omega = (R)omega_d, one_m = R(1) - omega;
  const bool relaxed = omega != R(1);
  auto upd = [&](T xold, T g) { return relaxed ? add(mul_real(xold, one_m), mul_real(g, omega)) : g; };
  if (nrows != ncols) return ORC_INCOMPATIBLE_FORMAT;  // :16-20 (GaussSeidel::new)
  if (!is_csr) return ORC_INCOMPATIBLE_FORMAT;         // :22-26
  if (n_rhs != nrows) return ORC_INCOMPATIBLE_FORMAT;  // :41-45
  if (n_rhs != n_x) return ORC_INCOMPATIBLE_FORMAT;    // :46-50
  if (max_iter == 0) {                                 // :52-54
    *iters = 0;
    return ORC_INSUFFICIENT_ITER;
  }
  const int64_t n = n_rhs;
  R b_norm = 0;
  T* res = ws;       // workspace[0..n]
  T* diag = ws + n;  // workspace[n..2n]
  for (int64_t row = 0; row < n; ++row) {  // :60-86 unrolled first sweep
    T sigma = zero_of(T());
    bool found = false;
    T d = zero_of(T());
    for (int64_t k = A.indptr[row]; k < A.indptr[row + 1]; ++k) {
      int64_t col = A.idx[k];
      if (row != col)
        sigma = add(sigma, mul(A.a[k], x[col]));  // :66
      else {
        d = A.a[k];  // :69
        found = true;
      }
    }
    if (!found || square(d) < EPS_T) {  // :72-78
      *iters = row;
      return ORC_ZERO_DIAGONAL;
    }
    diag[row] = d;                              // :81
    b_norm += square(rhs[row]);                 // :83
    x[row] = upd(x[row], divi(sub(rhs[row], sigma), d));  // :84
  }
  if (g_mode == 3) b_norm = exact_sumsq(n, rhs);  // exact-dot flavour: the same sum, rounded once
  const R tol2 = (R)eps * std::sqrt(b_norm);  // :87
  A.mul_vec(x, res);                            // :90
  axpy(n, neg(one_of(T())), rhs, res);          // :97
  R rn = norm2(n, res);                    // :104
  h.put(0, rn);
  if (rn <= tol2) {                             // :106-108
    *iters = 1;
    *resid = rn;
    return ORC_OK;
  }
  for (int64_t it = 1; it < max_iter; ++it) {   // :110
    for (int64_t row = 0; row < n; ++row) x[row] = upd(x[row], gs_row(A, row, rhs[row], diag[row], x));
    A.mul_vec(x, res);                          // :128
    axpy(n, neg(one_of(T())), rhs, res);        // :131
    rn = norm2(n, res);                         // :133
    h.put(it, rn);
    if (rn <= tol2) {                           // :135-137
      *iters = it;
      *resid = rn;
      return ORC_OK;
    }
  }
  *iters = max_iter;
  return ORC_INSUFFICIENT_ITER;  // :139
}

template <typename T>
int run_bicgstab(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                 const int32_t* idx, const void* a, int pc_kind, const void* pc_data,
                 const void* rhs, void* x, int64_t max_iter, double tol, void* work,
                 int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                 int64_t* hist_len) {
  Csr<T> A{size, indptr, idx, reinterpret_cast<const T*>(a)};
  Hist h{hist, hist_cap, 0};
  *iters = 0;
  *resid = 0.0;
  int st;
  if (pc_kind == ORC_PC_NONE) {
    st = bicgstab<T>(size, n_rhs, n_x, A, nullptr, reinterpret_cast<const T*>(rhs),
                     reinterpret_cast<T*>(x), max_iter, tol, reinterpret_cast<T*>(work), iters,
                     resid, h);
  } else {
    PrecondOwner<T> po(pc_kind, size, &A, pc_data);
    if (!po.ok) {
      *iters = po.bad_row;
      return ORC_ZERO_DIAGONAL;
    }
    st = bicgstab<T>(size, n_rhs, n_x, A, &po.p, reinterpret_cast<const T*>(rhs),
                     reinterpret_cast<T*>(x), max_iter, tol, reinterpret_cast<T*>(work), iters,
                     resid, h);
  }
  if (hist_len) *hist_len = h.len;
  return st;
}

template <typename T>
int run_minres(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr, const int32_t* idx,
               const void* a, int pc_kind, const void* pc_data, bool cs, const void* rhs,
               void* x, int64_t max_iter, double tol, void* work, int64_t* iters,
               double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) {
  Csr<T> A{size, indptr, idx, reinterpret_cast<const T*>(a)};
  Hist h{hist, hist_cap, 0};
  *iters = 0;
  *resid = 0.0;
  int st;
  if (pc_kind == ORC_PC_NONE) {
    st = minres<T>(size, n_rhs, n_x, A, nullptr, cs, reinterpret_cast<const T*>(rhs),
                   reinterpret_cast<T*>(x), max_iter, tol, reinterpret_cast<T*>(work), iters,
                   resid, h);
  } else {
    PrecondOwner<T> po(pc_kind, size, &A, pc_data);
    if (!po.ok) {
      *iters = po.bad_row;
      return ORC_ZERO_DIAGONAL;
    }
    st = minres<T>(size, n_rhs, n_x, A, &po.p, cs, reinterpret_cast<const T*>(rhs),
                   reinterpret_cast<T*>(x), max_iter, tol, reinterpret_cast<T*>(work), iters,
                   resid, h);
  }
  if (hist_len) *hist_len = h.len;
  return st;
}

template <typename T>
int run_gs(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
           const int64_t* indptr, const int32_t* idx, const void* a, const void* rhs,
           void* x, int64_t max_iter, double eps, void* work, int64_t* iters, double* resid,
           double* hist, int64_t hist_cap, int64_t* hist_len, double omega = 1.0) {
  Csr<T> A{nrows, indptr, idx, reinterpret_cast<const T*>(a)};
  Hist h{hist, hist_cap, 0};
  *iters = 0;
  *resid = 0.0;
  int st = gauss_seidel<T>(nrows, ncols, is_csr, n_rhs, n_x, A, reinterpret_cast<const T*>(rhs),
                           reinterpret_cast<T*>(x), max_iter, eps, reinterpret_cast<T*>(work),
                           iters, resid, h, omega);
  if (hist_len) *hist_len = h.len;
  return st;
}

template <typename T>
int run_gs_apply(int64_t n, const int64_t* indptr, const int32_t* idx, const void* a,
                 int symmetric, const void* in, void* out, double omega = 1.0) {
  Csr<T> A{n, indptr, idx, reinterpret_cast<const T*>(a)};
  const real_t<T> w = (real_t<T>)omega;
  PrecondOwner<T> po(symmetric ? ORC_PC_SSOR : ORC_PC_SOR_FWD, n, &A, &w);
  if (!po.ok) return ORC_ZERO_DIAGONAL;
  po.p.apply(reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out));
  return ORC_OK;
}

template <typename F>
int64_t gen_lap3d7(int64_t nx, int64_t ny, int64_t nz, int64_t* indptr, int32_t* idx,
                          F put_val) {
  const int64_t n = nx * ny * nz;
  if (!indptr) {  // closed-form count
    return 7 * n - 2 * (nx * ny + ny * nz + nx * nz);
  }
  // pass 1: row counts -> indptr (parallel-friendly two-pass)
#pragma omp parallel for schedule(static)
  for (int64_t row = 0; row < n; ++row) {
    int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
    int c = 1 + (x > 0) + (x + 1 < nx) + (y > 0) + (y + 1 < ny) + (z > 0) + (z + 1 < nz);
    indptr[row + 1] = c;
  }
  indptr[0] = 0;
  for (int64_t row = 0; row < n; ++row) indptr[row + 1] += indptr[row];
  if (idx) {
#pragma omp parallel for schedule(static)
    for (int64_t row = 0; row < n; ++row) {
      int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
      int64_t k = indptr[row];
      if (z > 0) { idx[k] = (int32_t)(row - nx * ny); put_val(k, false); ++k; }
      if (y > 0) { idx[k] = (int32_t)(row - nx); put_val(k, false); ++k; }
      if (x > 0) { idx[k] = (int32_t)(row - 1); put_val(k, false); ++k; }
      idx[k] = (int32_t)row; put_val(k, true); ++k;
      if (x + 1 < nx) { idx[k] = (int32_t)(row + 1); put_val(k, false); ++k; }
      if (y + 1 < ny) { idx[k] = (int32_t)(row + nx); put_val(k, false); ++k; }
      if (z + 1 < nz) { idx[k] = (int32_t)(row + nx * ny); put_val(k, false); ++k; }
    }
  }
  return indptr[n];
}


}  // namespace

extern "C" {

void orc_set_mode(int m) { g_mode = m; }
int orc_get_mode(void) { return g_mode; }
int orc_max_threads(void) { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }

void orc_spmv_d(int64_t n, const int64_t* ip, const int32_t* idx, const double* a, const double* x,
                double* y) {
  spmv_serial<double>(n, ip, idx, a, x, y);
}
void orc_spmv_z(int64_t n, const int64_t* ip, const int32_t* idx, const double* a, const double* x,
                double* y) {
  spmv_serial<cplx>(n, ip, idx, (const cplx*)a, (const cplx*)x, (cplx*)y);
}
void orc_spmv_par_d(int64_t n, const int64_t* ip, const int32_t* idx, const double* a,
                    const double* x, double* y) {
  spmv_par<double>(n, ip, idx, a, x, y);
}
void orc_spmv_par_z(int64_t n, const int64_t* ip, const int32_t* idx, const double* a,
                    const double* x, double* y) {
  spmv_par<cplx>(n, ip, idx, (const cplx*)a, (const cplx*)x, (cplx*)y);
}
// mat.rs:130-142: CSC scatter, y zero-filled first (mat.rs:71)
void orc_spmv_csc_d(int64_t nrows, int64_t ncols, const int64_t* ip, const int32_t* idx,
                    const double* a, const double* x, double* y) {
  for (int64_t i = 0; i < nrows; ++i) y[i] = 0.0;
  for (int64_t c = 0; c < ncols; ++c) {
    const double m = x[c];
    for (int64_t k = ip[c]; k < ip[c + 1]; ++k) y[idx[k]] += m * a[k];
  }
}
void orc_spmv_csc_z(int64_t nrows, int64_t ncols, const int64_t* ip, const int32_t* idx,
                    const double* a_, const double* x_, double* y_) {
  const cplx* a = (const cplx*)a_;
  const cplx* x = (const cplx*)x_;
  cplx* y = (cplx*)y_;
  for (int64_t i = 0; i < nrows; ++i) y[i] = cplx{0.0, 0.0};
  for (int64_t c = 0; c < ncols; ++c) {
    const cplx m = x[c];
    for (int64_t k = ip[c]; k < ip[c + 1]; ++k) y[idx[k]] = add(y[idx[k]], mul(m, a[k]));  // *t += *multiplier * value
  }
}
void orc_spmv_dot_d(int64_t n, const int64_t* ip, const int32_t* idx, const double* a,
                    const double* x, double* y, double* out) {
  Csr<double> A{n, ip, idx, a};
  int m = g_mode;
  g_mode = 0;
  *out = A.mul_vec_dot(x, y);
  g_mode = m;
}
void orc_spmv_dot_z(int64_t n, const int64_t* ip, const int32_t* idx, const double* a,
                    const double* x, double* y, double* out) {
  Csr<cplx> A{n, ip, idx, (const cplx*)a};
  int m = g_mode;
  g_mode = 0;
  cplx r = A.mul_vec_dot((const cplx*)x, (cplx*)y);
  g_mode = m;
  out[0] = r.re;
  out[1] = r.im;
}

double orc_dot_d(int64_t n, const double* x, const double* y) { return dot_fb<double>(n, x, y); }
double orc_conj_dot_d(int64_t n, const double* x, const double* y) {
  return conj_dot<double>(n, x, y);
}
double orc_norm2_d(int64_t n, const double* x) { return norm2<double>(n, x); }
void orc_dot_z(int64_t n, const double* x, const double* y, double* out) {
  cplx r = dot_fb<cplx>(n, (const cplx*)x, (const cplx*)y);
  out[0] = r.re;
  out[1] = r.im;
}
void orc_conj_dot_z(int64_t n, const double* x, const double* y, double* out) {
  cplx r = conj_dot<cplx>(n, (const cplx*)x, (const cplx*)y);
  out[0] = r.re;
  out[1] = r.im;
}
double orc_norm2_z(int64_t n, const double* x) { return norm2<cplx>(n, (const cplx*)x); }
void orc_axpy_d(int64_t n, double a, const double* x, double* y) { axpy<double>(n, a, x, y); }
void orc_axpby_d(int64_t n, double a, const double* x, double b, double* y) {
  axpby<double>(n, a, x, b, y);
}
void orc_scale_d(int64_t n, double a, double* x) { scale<double>(n, a, x); }
void orc_axpy_z(int64_t n, const double* a, const double* x, double* y) {
  axpy<cplx>(n, cplx{a[0], a[1]}, (const cplx*)x, (cplx*)y);
}
void orc_axpby_z(int64_t n, const double* a, const double* x, const double* b, double* y) {
  axpby<cplx>(n, cplx{a[0], a[1]}, (const cplx*)x, cplx{b[0], b[1]}, (cplx*)y);
}
void orc_scale_z(int64_t n, const double* a, double* x) {
  scale<cplx>(n, cplx{a[0], a[1]}, (cplx*)x);
}
void orc_rscale_z(int64_t n, double a, double* x) { rscale<cplx>(n, a, (cplx*)x); }
void orc_conj_z(int64_t n, const double* x, double* out) {
  conj_vec<cplx>(n, (const cplx*)x, (cplx*)out);
}

void orc_axpy_s(int64_t n, float a, const float* x, float* y) {
  for (int64_t i = 0; i < n; ++i) y[i] = y[i] + x[i] * a;
}
void orc_axpby_s(int64_t n, float a, const float* x, float b, float* y) {
  for (int64_t i = 0; i < n; ++i) y[i] = x[i] * a + y[i] * b;
}
float orc_conj_dot_s(int64_t n, const float* x, const float* y) { return conj_dot<float>(n, x, y); }
void orc_dot_c(int64_t n, const float* x, const float* y, float* out) {
  const cplxf r = dot_fb<cplxf>(n, (const cplxf*)x, (const cplxf*)y);
  out[0] = r.re;
  out[1] = r.im;
}
void orc_conj_dot_c(int64_t n, const float* x, const float* y, float* out) {
  const cplxf r = conj_dot<cplxf>(n, (const cplxf*)x, (const cplxf*)y);
  out[0] = r.re;
  out[1] = r.im;
}

void orc_diag_apply_d(int64_t n, const double* diag, const double* in, double* out) {
  PrecondOwner<double> po(ORC_PC_DIAG, n, nullptr, diag);
  po.p.apply(in, out);
}
void orc_diag_apply_z(int64_t n, const double* diag, const double* in, double* out) {
  PrecondOwner<cplx> po(ORC_PC_DIAG, n, nullptr, diag);
  po.p.apply((const cplx*)in, (cplx*)out);
}
void orc_diag_apply_zd(int64_t n, const double* diag, const double* in, double* out) {
  PrecondOwner<cplx> po(ORC_PC_DIAG_REAL, n, nullptr, diag);
  po.p.apply((const cplx*)in, (cplx*)out);
}
int orc_gs_apply_d(int64_t n, const int64_t* ip, const int32_t* idx, const double* a, int sym,
                   const double* in, double* out) {
  return run_gs_apply<double>(n, ip, idx, a, sym, in, out);
}
int orc_gs_apply_z(int64_t n, const int64_t* ip, const int32_t* idx, const double* a, int sym,
                   const double* in, double* out) {
  return run_gs_apply<cplx>(n, ip, idx, a, sym, in, out);
}

#define SOLVER_ARGS                                                                              \
  int64_t size, int64_t n_rhs, int64_t n_x, const int64_t *indptr, const int32_t *idx,          \
      const double *a, int pc_kind, const double *pc_data, const double *rhs, double *x,        \
      int64_t max_iter, double tol, double *work, int64_t *iters, double *resid, double *hist,  \
      int64_t hist_cap, int64_t *hist_len

int orc_bicgstab_d(SOLVER_ARGS) {
  return run_bicgstab<double>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, rhs, x, max_iter,
                              tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_bicgstab_z(SOLVER_ARGS) {
  return run_bicgstab<cplx>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, rhs, x, max_iter,
                            tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_minres_d(SOLVER_ARGS) {
  return run_minres<double>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, false, rhs, x,
                            max_iter, tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_minres_z(SOLVER_ARGS) {
  return run_minres<cplx>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, false, rhs, x,
                          max_iter, tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_csminres_d(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, const double* rhs, double* x,
                   int64_t max_iter, double tol, double* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_minres<double>(size, n_rhs, n_x, indptr, idx, a, ORC_PC_NONE, nullptr, true, rhs, x,
                            max_iter, tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_csminres_z(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, const double* rhs, double* x,
                   int64_t max_iter, double tol, double* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_minres<cplx>(size, n_rhs, n_x, indptr, idx, a, ORC_PC_NONE, nullptr, true, rhs, x,
                          max_iter, tol, work, iters, resid, hist, hist_cap, hist_len);
}
int orc_gauss_seidel_d(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
                       const int64_t* indptr, const int32_t* idx, const double* a,
                       const double* rhs, double* x, int64_t max_iter, double eps, double* work,
                       int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                       int64_t* hist_len) {
  return run_gs<double>(nrows, ncols, is_csr, n_rhs, n_x, indptr, idx, a, rhs, x, max_iter, eps,
                        work, iters, resid, hist, hist_cap, hist_len);
}
int orc_gauss_seidel_z(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
                       const int64_t* indptr, const int32_t* idx, const double* a,
                       const double* rhs, double* x, int64_t max_iter, double eps, double* work,
                       int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                       int64_t* hist_len) {
  return run_gs<cplx>(nrows, ncols, is_csr, n_rhs, n_x, indptr, idx, a, rhs, x, max_iter, eps,
                      work, iters, resid, hist, hist_cap, hist_len);
}

// ---------- f32 / Complex32 entry points (float arrays, interleaved re,im for _c) -------------
#define SOLVER_ARGS_F                                                                            \
  int64_t size, int64_t n_rhs, int64_t n_x, const int64_t *indptr, const int32_t *idx,          \
      const float *a, int pc_kind, const float *pc_data, const float *rhs, float *x,            \
      int64_t max_iter, double tol, float *work, int64_t *iters, double *resid, double *hist,   \
      int64_t hist_cap, int64_t *hist_len
void orc_spmv_s(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, const float* x, float* y) {
  spmv_serial<float>(n, ip, idx, a, x, y);
}
void orc_spmv_c(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, const float* x, float* y) {
  spmv_serial<cplxf>(n, ip, idx, (const cplxf*)a, (const cplxf*)x, (cplxf*)y);
}
float orc_spmv_dot_s(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, const float* x, float* y) {
  Csr<float> A{n, ip, idx, a};
  int m = g_mode;
  g_mode = 0;
  const float r = A.mul_vec_dot(x, y);
  g_mode = m;
  return r;
}
void orc_spmv_dot_c(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, const float* x, float* y, float* out) {
  Csr<cplxf> A{n, ip, idx, (const cplxf*)a};
  int m = g_mode;
  g_mode = 0;
  const cplxf r = A.mul_vec_dot((const cplxf*)x, (cplxf*)y);
  g_mode = m;
  out[0] = r.re;
  out[1] = r.im;
}
float orc_norm2_s(int64_t n, const float* x) { return norm2<float>(n, x); }
float orc_norm2_c(int64_t n, const float* x) { return norm2<cplxf>(n, (const cplxf*)x); }
float orc_dot_s(int64_t n, const float* x, const float* y) { return dot_fb<float>(n, x, y); }
void orc_scale_s(int64_t n, float a, float* x) { scale<float>(n, a, x); }
void orc_scale_c(int64_t n, const float* a, float* x) { scale<cplxf>(n, cplxf{a[0], a[1]}, (cplxf*)x); }
void orc_rscale_s(int64_t n, float a, float* x) { rscale<float>(n, a, x); }
void orc_rscale_c(int64_t n, float a, float* x) { rscale<cplxf>(n, a, (cplxf*)x); }
void orc_axpy_c(int64_t n, const float* a, const float* x, float* y) {
  axpy<cplxf>(n, cplxf{a[0], a[1]}, (const cplxf*)x, (cplxf*)y);
}
void orc_axpby_c(int64_t n, const float* a, const float* x, const float* b, float* y) {
  axpby<cplxf>(n, cplxf{a[0], a[1]}, (const cplxf*)x, cplxf{b[0], b[1]}, (cplxf*)y);
}
void orc_conj_c(int64_t n, const float* x, float* out) { conj_vec<cplxf>(n, (const cplxf*)x, (cplxf*)out); }
void orc_diag_apply_s(int64_t n, const float* diag, const float* in, float* out) {
  PrecondOwner<float> po(ORC_PC_DIAG, n, nullptr, diag);
  po.p.apply(in, out);
}
void orc_diag_apply_c(int64_t n, const float* diag, const float* in, float* out) {
  PrecondOwner<cplxf> po(ORC_PC_DIAG, n, nullptr, diag);
  po.p.apply((const cplxf*)in, (cplxf*)out);
}
void orc_diag_apply_cs(int64_t n, const float* diag, const float* in, float* out) {  // DiagPrecond<Complex32, f32>
  PrecondOwner<cplxf> po(ORC_PC_DIAG_REAL, n, nullptr, diag);
  po.p.apply((const cplxf*)in, (cplxf*)out);
}
int orc_gs_apply_s(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, int sym, const float* in, float* out) {
  return run_gs_apply<float>(n, ip, idx, a, sym, in, out);
}
int orc_gs_apply_c(int64_t n, const int64_t* ip, const int32_t* idx, const float* a, int sym, const float* in, float* out) {
  return run_gs_apply<cplxf>(n, ip, idx, a, sym, in, out);
}
int orc_bicgstab_s(SOLVER_ARGS_F) {
  return run_bicgstab<float>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, rhs, x, max_iter, tol, work, iters, resid,
                             hist, hist_cap, hist_len);
}
int orc_bicgstab_c(SOLVER_ARGS_F) {
  return run_bicgstab<cplxf>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, rhs, x, max_iter, tol, work, iters, resid,
                             hist, hist_cap, hist_len);
}
int orc_minres_s(SOLVER_ARGS_F) {
  return run_minres<float>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, false, rhs, x, max_iter, tol, work, iters,
                           resid, hist, hist_cap, hist_len);
}
int orc_minres_c(SOLVER_ARGS_F) {
  return run_minres<cplxf>(size, n_rhs, n_x, indptr, idx, a, pc_kind, pc_data, false, rhs, x, max_iter, tol, work, iters,
                           resid, hist, hist_cap, hist_len);
}
int orc_csminres_s(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr, const int32_t* idx, const float* a,
                   const float* rhs, float* x, int64_t max_iter, double tol, float* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_minres<float>(size, n_rhs, n_x, indptr, idx, a, ORC_PC_NONE, nullptr, true, rhs, x, max_iter, tol, work, iters,
                           resid, hist, hist_cap, hist_len);
}
int orc_csminres_c(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr, const int32_t* idx, const float* a,
                   const float* rhs, float* x, int64_t max_iter, double tol, float* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_minres<cplxf>(size, n_rhs, n_x, indptr, idx, a, ORC_PC_NONE, nullptr, true, rhs, x, max_iter, tol, work, iters,
                           resid, hist, hist_cap, hist_len);
}
int orc_gauss_seidel_s(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                       const int32_t* idx, const float* a, const float* rhs, float* x, int64_t max_iter, double eps,
                       float* work, int64_t* iters, double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_gs<float>(nrows, ncols, is_csr, n_rhs, n_x, indptr, idx, a, rhs, x, max_iter, eps, work, iters, resid, hist,
                       hist_cap, hist_len);
}
int orc_gauss_seidel_c(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                       const int32_t* idx, const float* a, const float* rhs, float* x, int64_t max_iter, double eps,
                       float* work, int64_t* iters, double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) {
  return run_gs<cplxf>(nrows, ncols, is_csr, n_rhs, n_x, indptr, idx, a, rhs, x, max_iter, eps, work, iters, resid, hist,
                       hist_cap, hist_len);
}

// ---------- relaxed Gauss-Seidel (SOR / SSOR(w)): build-side additions, see sprs_oracle.h ----------
#define ORC_SOR_APPLY(sfx, T, F)                                                                                     \
  int orc_sor_apply_##sfx(int64_t n, const int64_t* ip, const int32_t* idx, const F* a, int sym, double omega,      \
                          const F* in, F* out) {                                                                    \
    return run_gs_apply<T>(n, ip, idx, a, sym, in, out, omega);                                                     \
  }                                                                                                                  \
  int orc_sor_solve_##sfx(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x, const int64_t* ip,   \
                          const int32_t* idx, const F* a, const F* rhs, F* x, int64_t max_iter, double eps,          \
                          double omega, F* work, int64_t* iters, double* resid, double* hist, int64_t hist_cap,      \
                          int64_t* hist_len) {                                                                       \
    return run_gs<T>(nrows, ncols, is_csr, n_rhs, n_x, ip, idx, a, rhs, x, max_iter, eps, work, iters, resid, hist,  \
                     hist_cap, hist_len, omega);                                                                     \
  }
ORC_SOR_APPLY(d, double, double)
ORC_SOR_APPLY(z, cplx, double)
ORC_SOR_APPLY(s, float, float)
ORC_SOR_APPLY(c, cplxf, float)

// ---------- generators ----------------------------------------------------------------------
static inline bool is_border(int64_t r, int64_t c, int64_t rows, int64_t cols) {
  return r == 0 || r + 1 == rows || c == 0 || c + 1 == cols;  // main.rs:40-51
}

int64_t orc_gen_dirichlet2d(int64_t rows, int64_t cols, int64_t* indptr, int32_t* idx, double* a,
                            double* rhs) {
  // main.rs:53-88; column index i*rows+j as written in the reference (square grids).
  int64_t cum = 0;
  for (int64_t i = 0; i < rows; ++i)
    for (int64_t j = 0; j < cols; ++j) {
      const int64_t row = i * cols + j;
      if (indptr) indptr[row] = cum;
      auto put = [&](int64_t ii, int64_t jj, double v) {
        if (idx) {
          idx[cum] = (int32_t)(ii * rows + jj);
          a[cum] = v;
        }
        ++cum;
      };
      if (is_border(i, j, rows, cols)) {
        put(i, j, 1.0);
        if (rhs) rhs[i * rows + j] = (double)(i + j);  // main.rs:9-11, 90-103
      } else {
        put(i - 1, j, 1.0);
        put(i, j - 1, 1.0);
        put(i, j, -4.0);
        put(i, j + 1, 1.0);
        put(i + 1, j, 1.0);
        if (rhs) rhs[i * rows + j] = 0.0;
      }
    }
  if (indptr) indptr[rows * cols] = cum;
  return cum;
}

int64_t orc_gen_lap3d7_d(int64_t nx, int64_t ny, int64_t nz, double shift, int64_t* indptr,
                         int32_t* idx, double* a) {
  const double diag = 6.0 - shift;
  return gen_lap3d7(nx, ny, nz, indptr, idx,
                    [=](int64_t k, bool d) { a[k] = d ? diag : -1.0; });
}
int64_t orc_gen_lap3d7_z(int64_t nx, int64_t ny, int64_t nz, double sre, double sim,
                         int64_t* indptr, int32_t* idx, double* a) {
  const double dre = 6.0 - sre, dim = 0.0 - sim;
  return gen_lap3d7(nx, ny, nz, indptr, idx, [=](int64_t k, bool d) {
    a[2 * k] = d ? dre : -1.0;
    a[2 * k + 1] = d ? dim : 0.0;
  });
}

int64_t orc_gen_convdiff27_d(int64_t nx, int64_t ny, int64_t nz, double bx, double by, double bz,
                             int64_t row_begin, int64_t row_end, int64_t* indptr, int32_t* idx,
                             double* a) {
  const int64_t nloc = row_end - row_begin;
  auto count_row = [&](int64_t row) {
    int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
    int cx = 1 + (x > 0) + (x + 1 < nx), cy = 1 + (y > 0) + (y + 1 < ny),
        cz = 1 + (z > 0) + (z + 1 < nz);
    return (int64_t)cx * cy * cz;
  };
  if (!indptr) {
    int64_t tot = 0;
#pragma omp parallel for reduction(+ : tot) schedule(static)
    for (int64_t row = row_begin; row < row_end; ++row) tot += count_row(row);
    return tot;
  }
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nloc; ++r) indptr[r + 1] = count_row(row_begin + r);
  indptr[0] = 0;
  for (int64_t r = 0; r < nloc; ++r) indptr[r + 1] += indptr[r];
  if (idx) {
    const double centre = 26.0 + bx + by + bz;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nloc; ++r) {
      const int64_t row = row_begin + r;
      int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
      int64_t k = indptr[r];
      for (int dz = -1; dz <= 1; ++dz) {
        if (z + dz < 0 || z + dz >= nz) continue;
        for (int dy = -1; dy <= 1; ++dy) {
          if (y + dy < 0 || y + dy >= ny) continue;
          for (int dx = -1; dx <= 1; ++dx) {
            if (x + dx < 0 || x + dx >= nx) continue;
            idx[k] = (int32_t)(row + dx + dy * nx + dz * nx * ny);
            double v;
            if (dx == 0 && dy == 0 && dz == 0)
              v = centre;
            else if (dx == -1 && dy == 0 && dz == 0)
              v = -1.0 - bx;
            else if (dx == 0 && dy == -1 && dz == 0)
              v = -1.0 - by;
            else if (dx == 0 && dy == 0 && dz == -1)
              v = -1.0 - bz;
            else
              v = -1.0;
            a[k] = v;
            ++k;
          }
        }
      }
    }
  }
  return indptr[nloc];
}

}  // extern "C"
