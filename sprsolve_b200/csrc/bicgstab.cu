// bicgstab.cu -- device-resident BiCGStab: BiCGStab::solve (src/bicg_stab.rs:35-200) and
// BiCGStab::precond_solve (src/bicg_stab.rs:204-366).
//
// Same recurrence, same sign convention (r = A x - b, x -= ...), same unrolled first iteration,
// rho restart, breakdown test and convergence test at the top of the next iteration.  The
// reference's 11 separate vecalg passes per iteration are fused into three kernels and two SpMV
// epilogues; every element is still computed with the reference's operations in the reference's
// order (no FMA), only the long sums are re-ordered:
//   S1 : ||r||, rho = <r0,r>  -> convergence / restart test, beta                (:296-319)
//   K1 : p = (-beta w) v + beta p ; p += r ; y = M p        [Jacobi fused]        (:324-328)
//   SpMV1: v = A y, epilogue <r0,v>                                               (:329-332)
//   S2 : breakdown test, alpha = rho / <r0,v>                                     (:333-338)
//   K2 : r -= alpha v ; z = M r                             [Jacobi fused]        (:341-343)
//   SpMV2: t = A z, epilogue <t,t>, <t,r>                                         (:344-349)
//   S3 : w = <t,r>/<t,t> (0 if <t,t>.re <= 0)                                     (:347-352)
//   K3 : x -= alpha y ; x -= w z ; r -= w t ; partials of ||r||^2 and <r0,r>      (:355-362,296,301)
// n-vector streams per iteration: K1 4R+2W, SpMV1 epilogue 1R, K2 3R+2W, SpMV2 epilogue 1R,
// K3 6R+2W = 21 (Jacobi); 16 without preconditioner.
#include <algorithm>
#include <vector>

#include "finalize.cuh"
#include "solver.cuh"

namespace spb {

template <typename T>
struct BicgState {
  StateHead h;
  T rho, rho_old, alpha, w, beta, c_pv, nalpha, nw;
  real_t<T> rhs_norm, tol2, r0_norm_tol, r_norm, tol;  // T::Real quantities (f32 solvers: float, as in the reference)
};

__device__ __forceinline__ void hist_put(StateHead& h, double* hist, long long cap, long long k, double v) {
  if (hist && k < cap) hist[k] = v;
  if (k + 1 > h.hist_len) h.hist_len = k + 1;
}

// ---------------------------------------------------------------- scalar kernels (1 thread)
template <typename T>
__device__ __forceinline__ void bicg_s_rhs_body(BicgState<T>* st, const scal2* red) {
  using R = real_t<T>;
  const R rhs_norm = sqrt_r((R)red[0].re);  // norm2(rhs), :225
  st->rhs_norm = rhs_norm;
  st->tol2 = st->tol * rhs_norm;            // :231
  if (rhs_norm <= eps_of<T>()) {            // :226-230
    st->h.status = DS_ZERO_RHS;
    st->h.res_iters = 0;
    st->h.res_resid = (double)rhs_norm;
  }
}

template <typename T>
__global__ void bicg_s_rhs(BicgState<T>* st, const scal2* red) {
  bicg_s_rhs_body(st, red);
}

// after r = A x - rhs, r0 = r: used for the start (:251-259) and for the restart (:315-317)
template <typename T>
__device__ __forceinline__ void bicg_s_init_body(BicgState<T>* st, const scal2* red, double* hist, long long cap, int restart) {
  if (restart ? st->h.status != DS_NEED_RESTART : st->h.status != DS_RUNNING) return;
  using R = real_t<T>;
  const R rn = sqrt_r((R)red[0].re);
  if (!restart) {
    hist_put(st->h, hist, cap, 0, (double)(rn / st->rhs_norm));
    if (rn <= st->tol2) {  // :252-254
      st->h.status = DS_OK;
      st->h.res_iters = 0;
      st->h.res_resid = (double)(rn / st->rhs_norm);
      return;
    }
    R t = rn * eps_of<T>();  // :255-256
    st->r0_norm_tol = t * t;
    st->rho = from_real<T>(rn * rn);  // :259
  } else {
    st->rho = from_real<T>(rn * rn);                                        // :316
    st->r0_norm_tol = re_of(st->rho) * eps_of<T>() * eps_of<T>();         // :317
    const T beta = mul(divi(st->rho, st->rho_old), divi(st->alpha, st->w));  // :319
    st->beta = beta;
    st->c_pv = mul(neg(beta), st->w);
    st->h.status = DS_RUNNING;
  }
}

template <typename T>
__global__ void bicg_s_init(BicgState<T>* st, const scal2* red, double* hist, long long cap, int restart) {
  bicg_s_init_body(st, red, hist, cap, restart);
}

template <typename T>
__device__ __forceinline__ void bicg_s1_body(BicgState<T>* st, const scal2* red, double* hist, long long cap) {
  if (st->h.status != DS_RUNNING) return;
  const long long its = ++st->h.its;
  using R = real_t<T>;
  const R r_norm = sqrt_r((R)red[0].re);  // :296
  st->r_norm = r_norm;
  hist_put(st->h, hist, cap, its, (double)(r_norm / st->rhs_norm));
  if (r_norm <= st->tol2) {  // :297-299
    st->h.status = DS_OK;
    st->h.res_iters = its;
    st->h.res_resid = (double)(r_norm / st->rhs_norm);
    return;
  }
  st->rho_old = st->rho;             // :300
  st->rho = from_scal2<T>(red[1]);   // :301
  if (abs_of(st->rho) < st->r0_norm_tol) {  // :304
    st->h.status = DS_NEED_RESTART;
    return;
  }
  const T beta = mul(divi(st->rho, st->rho_old), divi(st->alpha, st->w));  // :319
  st->beta = beta;
  st->c_pv = mul(neg(beta), st->w);  // -beta * w, :324
}

template <typename T>
__global__ void bicg_s1(BicgState<T>* st, const scal2* red, double* hist, long long cap) {
  bicg_s1_body(st, red, hist, cap);
}

template <typename T>
__device__ __forceinline__ void bicg_s2_body(BicgState<T>* st, const scal2* red, int first) {
  if (st->h.status != DS_RUNNING) return;
  const T tmp = from_scal2<T>(red[0]);  // conj_dot(r0, v), :332
  if (!first && abs_of(tmp) <= (real_t<T>)0) {   // :333-336 (the unrolled first iteration has no test, :266)
    st->h.status = DS_BREAKDOWN;
    st->h.res_iters = st->h.its;
    return;
  }
  st->alpha = divi(st->rho, tmp);  // :338
  st->nalpha = neg(st->alpha);
}
// fused into the kernel that finishes <r0, v> (finalize.cuh)
template <typename T>
struct BicgS2Tail {
  BicgState<T>* st;
  const scal2* red;
  int first;
  __device__ __forceinline__ void operator()() const { bicg_s2_body(st, red, first); }
};

template <typename T>
__device__ __forceinline__ void bicg_s3_body(BicgState<T>* st, const scal2* red) {
  if (st->h.status != DS_RUNNING) return;
  const T tt = from_scal2<T>(red[0]);  // conj_dot(t, t), :347
  st->w = re_of(tt) > (real_t<T>)0 ? divi(from_scal2<T>(red[1]), tt) : zero_of<T>();  // :348-352
  st->nw = neg(st->w);
}
// fused into the kernel that finishes <t, t>, <t, r>
template <typename T>
struct BicgS3Tail {
  BicgState<T>* st;
  const scal2* red;
  __device__ __forceinline__ void operator()() const { bicg_s3_body(st, red); }
};

// ---------------------------------------------------------------- vector kernels
template <typename T>
__global__ void __launch_bounds__(kVecThreads)
bicg_k_init(const BicgState<T>* st, int restart, int64_t n, const T* rhs, T* r, T* r0, Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>();
  if (restart ? st->h.status == DS_NEED_RESTART : st->h.status == DS_RUNNING) {
    const T m1 = neg(one_of<T>());
    SPB_GRID_STRIDE(i, n) {
      const T ri = add(r[i], mul(rhs[i], m1));  // axpy(-1, rhs, r), :246
      r[i] = ri;
      r0[i] = ri;                               // :249
      acc_sq(e0, ri);                           // norm2, :251
    }
  }
  write_partials(e0, zero_of<Acc<T>>(), partials);
}

template <typename T, typename V, bool FIRST, bool WRITE_Y>
__global__ void __launch_bounds__(kVecThreads)
bicg_k1(const BicgState<T>* st, int64_t n, const T* v, T* p, const T* r, T* y, const V* dinv) {
  if (st->h.status != DS_RUNNING) return;
  const T c_pv = st->c_pv, beta = st->beta, one = one_of<T>();
  SPB_GRID_STRIDE(i, n) {
    T pi;
    if (FIRST) {
      pi = r[i];  // p = r, :261
    } else {
      pi = add(mul(v[i], c_pv), mul(p[i], beta));  // axpby(-beta w, v, beta, p), :324
      pi = add(pi, mul(r[i], one));                // axpy(1, r, p), :325
    }
    p[i] = pi;
    if (WRITE_Y) y[i] = mul_diag(pi, dinv[i]);     // y = M p, :328 (src/precond.rs:50)
  }
}

template <typename T, typename V, bool WRITE_Z>
__global__ void __launch_bounds__(kVecThreads)
bicg_k2(const BicgState<T>* st, int64_t n, T* r, const T* v, T* z, const V* dinv) {
  if (st->h.status != DS_RUNNING) return;
  const T nalpha = st->nalpha;
  SPB_GRID_STRIDE(i, n) {
    const T ri = add(r[i], mul(v[i], nalpha));  // axpy(-alpha, v, r), :341
    r[i] = ri;
    if (WRITE_Z) z[i] = mul_diag(ri, dinv[i]);  // z = M r, :343
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads)
bicg_k3(const BicgState<T>* st, int64_t n, T* x, const T* y, const T* z, T* r, const T* t, const T* r0,
        Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
  if (st->h.status == DS_RUNNING) {
    const T nalpha = st->nalpha, nw = st->nw;
    SPB_GRID_STRIDE(i, n) {
      const T zi = z[i];  // may alias r (no preconditioner): read before r is updated
      T xi = add(x[i], mul(y[i], nalpha));  // axpy(-alpha, y, x), :355
      xi = add(xi, mul(zi, nw));            // axpy(-w, z, x), :357
      x[i] = xi;
      const T ri = add(r[i], mul(t[i], nw));  // axpy(-w, t, r), :362
      r[i] = ri;
      acc_sq(e0, ri);                    // next norm2(r), :296
      acc_prod(e1, conj_of(r0[i]), ri);  // next conj_dot(r0, r), :301
    }
  }
  write_partials(e0, e1, partials);
}

// ---------------------------------------------------------------- single-kernel solve (L2-resident systems)
// On a system whose matrix and vectors fit in L2 one iteration is a few microseconds of work and the
// multi-kernel loop above is bound by per-kernel latency (~5-12 us x 9.5 launches: 63 us on the 512^2
// reference matrix, independent of n).  Here the WHOLE solve -- ||b||, r = A x - b, the unrolled first
// iteration, the loop with its convergence / restart / breakdown tests, the residual history -- is ONE
// kernel: CTA b owns the contiguous rows [b R, (b+1) R), a thread owns the same rows in every phase, the
// vectors only their owner touches live in shared memory.  An iteration has five synchronisation points
// (three reductions, two "the gathered vector is complete"); what they cost decides the iteration time:
//   * MODE 0 / 1 (cooperative grid, one CTA per SM): a counter barrier -- red.release.gpu arrive, ONE thread
//     per CTA polling with relaxed loads (no acquire in the loop: an acquire invalidates the SM's L1 and with
//     it the matrix lines), the data read afterwards at L2 (ld.cg).  A reduction point stores the CTA partial
//     before its arrive and the first warp of EVERY CTA sums all CTA partials itself (fixed order, one
//     rounding): the Krylov scalars are replicated, no broadcast.  Measured on B200: polling MANY addresses
//     with relaxed (strong) loads costs ~10 clocks per load in the polling SM -- a data-as-flag exchange of
//     the 148 partials or of the halo entries was 1.5-2x slower than counter + plain loads -- so strong
//     accesses are kept to one flag per hand-off;
//   * MODE 2 (banded matrices, small systems: <= 16 CTAs = ONE thread-block cluster, launched with the
//     non-portable cluster size 16): nothing of an iteration leaves the SMs.  A CTA keeps the window
//     [row0 - bw_lo, row1 + bw_hi) of the gathered vector (y = M p, then z = M s) AND its slice of the matrix
//     in shared memory; the exchanges go through DISTRIBUTED SHARED MEMORY -- the hardware cluster barrier
//     instead of L2 round trips, reduction partials pushed into every CTA's inbox with st.shared::cluster,
//     halo entries read from the neighbours' windows with ld.shared::cluster.  This is the path of the
//     reference's own bench sizes (100^2 .. 140^2).
//   Also measured and dropped: the same window with a per-CTA release flag and L2 halo loads on the full grid
//   (three dependent L2 round trips per hand-off: slower than the counter barrier + L2 gathers, 30.7 vs 23.9 us
//   per iteration on 512^2) and register double-buffering of the next row's matrix entries (spills at 512 threads).
// Same element-wise operations and exactly rounded sums as the multi-kernel path => the same bits.
template <typename T>
struct FusedArgs {
  const int* indptr;
  const int* cols;
  const T* vals;
  int n;
  const T* rhs;
  T *x, *r, *r0, *p, *y, *v, *t, *z;
  const void* dinv;
  BicgState<T>* st;         // final state (written by CTA 0)
  double* rslots;           // [3][gridDim.x][2 * sizeof(Acc<T>) / 8] reduction slots (data-as-flag)
  unsigned long long* bar;  // grid barrier counter, zeroed before the launch
  double* hist;
  long long cap, max_iter;
  double tol;
  long long* stats;  // optional (SPB_FUSED_STATS=1): clocks of CTA 0 per phase, see the lap() calls
  int slots;             // shared-memory slots per own-row vector: rows_per_cta rounded up to a warp (not to a CTA: what
                         // shared memory does not take stays L1, where the matrix lines of the CTA's rows live)
  unsigned poll_sleep;   // ns between two polls of the grid barrier (0: spin)
  int rows_per_cta;      // R
  int bw_lo, bw_hi;      // max (row - col), max (col - row) over the matrix (MODE 2)
  int win_elems;         // capacity of the shared-memory window (MODE 2)
  int mat_cap;           // MODE 2: matrix entries of a CTA's rows that fit its shared memory (0: matrix stays in global memory)
};

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// ---- thread-block cluster / distributed shared memory (MODE 3)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local_smem, unsigned rank) {  // the same variable in CTA `rank`
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local_smem);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  return ra;
}
__device__ __forceinline__ void dsmem_st2(uint32_t ra, double a, double b) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1,%2};" ::"r"(ra), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double dsmem_ld(uint32_t ra, const double*) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ cplx dsmem_ld(uint32_t ra, const cplx*) {
  cplx v;
  asm volatile("ld.shared::cluster.v2.f64 {%0,%1}, [%2];" : "=d"(v.re), "=d"(v.im) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ float dsmem_ld(uint32_t ra, const float*) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ cplxf dsmem_ld(uint32_t ra, const cplxf*) {
  cplxf v;
  asm volatile("ld.shared::cluster.v2.f32 {%0,%1}, [%2];" : "=f"(v.re), "=f"(v.im) : "r"(ra) : "memory");
  return v;
}
// vectors other CTAs write during the kernel are read at L2 (the L1 of an SM is not coherent)
__device__ __forceinline__ double ld_l2(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_l2(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}
__device__ __forceinline__ void st_l2(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void st_l2(cplx* p, cplx v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.re, v.im)); }
__device__ __forceinline__ void st_l2(float* p, float v) { __stcg(p, v); }
__device__ __forceinline__ void st_l2(cplxf* p, cplxf v) { __stcg(reinterpret_cast<float2*>(p), make_float2(v.re, v.im)); }
__device__ __forceinline__ double ld_ro(const double* p) { return __ldg(p); }
__device__ __forceinline__ cplx ld_ro(const cplx* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}
__device__ __forceinline__ float ld_l2(const float* p) { return __ldcg(p); }
__device__ __forceinline__ cplxf ld_l2(const cplxf* p) {
  const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  return cplxf{v.x, v.y};
}
__device__ __forceinline__ float ld_ro(const float* p) { return __ldg(p); }
__device__ __forceinline__ cplxf ld_ro(const cplxf* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return cplxf{v.x, v.y};
}
template <typename T>
__device__ __forceinline__ scal2 acc_round(const AccR& a) {
  return scal2{round_to_real<T>(a.hi + a.lo), 0.0};
}
template <typename T>
__device__ __forceinline__ scal2 acc_round(const AccC& a) {
  return scal2{round_to_real<T>(a.rh + a.rl), round_to_real<T>(a.ih + a.il)};
}

// CTA partials of a reduction point in global memory (MODE 0-2): written before the arrive, read at L2 after it
__device__ __forceinline__ void slot_st(double* p, const AccR& a) { __stcg(reinterpret_cast<double2*>(p), make_double2(a.hi, a.lo)); }
__device__ __forceinline__ void slot_st(double* p, const AccC& a) {
  __stcg(reinterpret_cast<double2*>(p), make_double2(a.rh, a.rl));
  __stcg(reinterpret_cast<double2*>(p) + 1, make_double2(a.ih, a.il));
}
__device__ __forceinline__ void slot_ld(const double* p, AccR& a) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  a = AccR{v.x, v.y};
}
__device__ __forceinline__ void slot_ld(const double* p, AccC& a) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p)), w = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  a = AccC{v.x, v.y, w.x, w.y};
}
// the same through (distributed) shared memory (MODE 3)
__device__ __forceinline__ void dslot_st(uint32_t ra, const AccR& a) { dsmem_st2(ra, a.hi, a.lo); }
__device__ __forceinline__ void dslot_st(uint32_t ra, const AccC& a) {
  dsmem_st2(ra, a.rh, a.rl);
  dsmem_st2(ra + 16, a.ih, a.il);
}
__device__ __forceinline__ void sslot_ld(const double* p, AccR& a) { a = AccR{p[0], p[1]}; }
__device__ __forceinline__ void sslot_ld(const double* p, AccC& a) { a = AccC{p[0], p[1], p[2], p[3]}; }

// One CSR row folded sequentially in CSR order (src/mat.rs:100-105); matrix through the read-only path
// (constant for the whole kernel, L1-resident after the first iteration), x at L2.  Gathers are issued
// 8 per batch: a row of <= 8 entries costs one L2 round trip.
template <typename T>
__device__ __forceinline__ T fused_row(const FusedArgs<T>& a, int p0, int p1, const T* src) {
  T acc = zero_of<T>();
  for (int k = p0; k < p1; k += 8) {
    int c[8];
    T m[8], xv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = min(k + j, p1 - 1);
      c[j] = __ldg(a.cols + kk);
      m[j] = ld_ro(a.vals + kk);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = ld_l2(src + c[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (k + j < p1) acc = add(acc, mul(xv[j], m[j]));
  }
  return acc;
}
// the next own row's column / value lines into L1 while the current row waits for its gathers
template <typename T>
__device__ __forceinline__ void fused_prefetch_row(const FusedArgs<T>& a, int p0, int p1) {
  if (p1 > p0) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.cols + p0));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.vals + p0));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.vals + p1 - 1));
  }
}
__device__ __forceinline__ void red_release_gpu_add(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// sum of the first N (a power of two <= 32) lanes; valid in lane 0
template <int N, typename A>
__device__ __forceinline__ A warp_sum_n(A v) {
#pragma unroll
  for (int d = N / 2; d > 0; d >>= 1) v = add(v, shfl_down(v, d));
  return v;
}

__device__ __forceinline__ AccR shfl_idx(AccR v, int src) {
  return AccR{__shfl_sync(0xffffffffu, v.hi, src), __shfl_sync(0xffffffffu, v.lo, src)};
}
__device__ __forceinline__ AccC shfl_idx(AccC v, int src) {
  return AccC{__shfl_sync(0xffffffffu, v.rh, src), __shfl_sync(0xffffffffu, v.rl, src), __shfl_sync(0xffffffffu, v.ih, src),
              __shfl_sync(0xffffffffu, v.il, src)};
}

static const int kFusedMaxGrid = 160;  // slots a lane of the summing warp handles: kFusedMaxGrid / 32
static const int kFusedCluster = 16;   // MODE 2: CTAs of the one cluster (non-portable size, B200 allows 16)

// Cluster mode SpMV over the own rows: x from the shared-memory window, and the matrix slice of the CTA from
// shared memory too when it fits (columns stored window-relative) -- nothing in the row loop leaves the SM.
// (The cluster barrier flushes L1, so a matrix left in global memory would be re-read from L2 in every phase.)
template <typename T, int BLOCK, typename F>
__device__ __forceinline__ void fused_spmv_win(const FusedArgs<T>& a, const int2* ext, int nrows, int row0, const T* win, int wlo,
                                               const int* mcols, const T* mvals, int mbase, bool mat_smem, F&& emit) {
  for (int l = threadIdx.x; l < nrows; l += BLOCK) {
    const int2 e = ext[l];
    T acc = zero_of<T>();
    if (mat_smem) {
      // 8 entries per batch: the column -> x loads of a batch are independent, only the fold is a chain
      for (int k = e.x - mbase; k < e.y - mbase; k += 8) {
        const int k1 = e.y - mbase;
        T xv[8], m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kk = min(k + j, k1 - 1);
          xv[j] = win[mcols[kk]];
          m[j] = mvals[kk];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k + j < k1) acc = add(acc, mul(xv[j], m[j]));
      }
    } else {
      for (int k = e.x; k < e.y; k += 8) {
        int c[8];
        T m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kk = min(k + j, e.y - 1);
          c[j] = __ldg(a.cols + kk);
          m[j] = ld_ro(a.vals + kk);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (k + j < e.y) acc = add(acc, mul(win[c[j] - wlo], m[j]));
      }
    }
    emit(l, row0 + l, acc);
  }
}

// MODE 0: all vectors in global memory, counter barrier before each SpMV.
// MODE 1: the vectors only their owning thread touches (r, r0, p, v, t, x, 1/diag, row extents) live in
//         SHARED MEMORY for the whole solve (slot = local row), so the vector-update phases never wait for
//         L2; y and z go to global memory, counter barrier before each SpMV.
// MODE 2: MODE 1 inside ONE thread-block cluster (banded matrices, small systems): the window of the gathered
//         vector and the matrix slice in shared memory, all exchanges through distributed shared memory.
template <typename T, typename V, bool PC, int BLOCK, int MODE>
__global__ void __launch_bounds__(BLOCK, 1) bicg_fused_kernel(const FusedArgs<T> a) {
  constexpr bool SV = MODE >= 1, WIN = MODE == 2, CL = MODE == 2;
  constexpr int NW = BLOCK / 32;
  constexpr int AW = (int)sizeof(Acc<T>) / 8;  // doubles per accumulator
  constexpr int MAXS = kFusedMaxGrid / 32;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  __shared__ BicgState<T> S;
  __shared__ Acc<T> wsum[2][NW];
  __shared__ __align__(16) double inbox[2][kFusedCluster][2 * AW];  // MODE 2: the CTA partials of a reduction point
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = a.n, G = (int)gridDim.x;
  const int R = a.rows_per_cta;
  const int row0 = min(n, (int)blockIdx.x * R), row1 = min(n, row0 + R);
  const int nrows = row1 - row0;
  const V* dinv_g = static_cast<const V*>(a.dinv);
  double* hist = blockIdx.x == 0 ? a.hist : nullptr;  // one writer
  unsigned long long target = 0;  // arrivals the barrier counter must have seen at the next sync point
  int rp = 0;  // reduction points cycle through three slot buffers (two inboxes in MODE 3)
  scal2 red[2];
  // own-row vectors: shared memory (slot = local row) or global (index = row)
  const int slots = a.slots;
  T* const sm = reinterpret_cast<T*>(fused_smem);
  T* const r_v = SV ? sm : a.r;
  T* const r0_v = SV ? sm + slots : a.r0;
  T* const p_v = SV ? sm + 2 * slots : a.p;
  T* const v_v = SV ? sm + 3 * slots : a.v;
  T* const t_v = SV ? sm + 4 * slots : a.t;
  T* const x_v = SV ? sm + 5 * slots : a.x;
  const size_t off_d = (size_t)slots * 6 * sizeof(T);
  const size_t off_e = off_d + (PC ? (size_t)slots * sizeof(V) : 0);
  const size_t off_w = (off_e + (size_t)slots * sizeof(int2) + 15) & ~(size_t)15;
  const V* const d_v = SV ? reinterpret_cast<const V*>(fused_smem + off_d) : dinv_g;
  const int2* const ext_v = reinterpret_cast<const int2*>(fused_smem + off_e);
  // window of the gathered vector: columns [wlo, whi)
  const int wlo = WIN ? max(0, row0 - a.bw_lo) : 0, whi = WIN ? min(n, row1 + a.bw_hi) : 0;
  T* const win = reinterpret_cast<T*>(fused_smem + off_w);
  const int nlo = row0 - wlo, nhi = whi - row1;  // halo entries below / above the own rows
  // matrix slice of the own rows in shared memory (cluster mode, when it fits): columns window-relative
  const size_t off_mc = (off_w + (size_t)a.win_elems * sizeof(T) + 15) & ~(size_t)15;
  const size_t off_mv = (off_mc + (size_t)a.mat_cap * sizeof(int) + 15) & ~(size_t)15;
  int* const mcols = reinterpret_cast<int*>(fused_smem + off_mc);
  T* const mvals = reinterpret_cast<T*>(fused_smem + off_mv);
  const int mbase = WIN ? __ldg(a.indptr + row0) : 0;
  const bool mat_smem = WIN && a.mat_cap > 0 && (__ldg(a.indptr + row1) - mbase) <= a.mat_cap;
  auto extent = [&](int l, int i) -> int2 {
    if (SV) return ext_v[l];
    return make_int2(__ldg(a.indptr + i), __ldg(a.indptr + i + 1));
  };
#define SPB_OWN(l, i) (SV ? (l) : (i))
#define SPB_ROWS(l, i) for (int l = tid, i = row0 + tid; l < nrows; l += BLOCK, i += BLOCK)
  long long t_last = a.stats ? clock64() : 0;
  long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  auto lap = [&](int k) {  // diagnostics: phase clocks per CTA (accumulated in registers, stored at the end)
    if (a.stats && tid == 0) {
      const long long now = clock64();
      t_acc[k] += now - t_last;
      t_last = now;
    }
  };

  // Barrier over all CTAs.  Grid: a release-add without return value and a relaxed spin by one thread per CTA
  // (what is read afterwards is read at L2); __syncthreads on both sides extends the ordering to the CTA.
  // Cluster: the hardware barrier (release / acquire at cluster scope, shared and global memory).
  auto arrive_and_wait = [&]() {  // thread 0 only
    target += gridDim.x;
    red_release_gpu_add(a.bar, 1ULL);
    while (ld_relaxed_gpu_u64(a.bar) < target) {
      if (a.poll_sleep) __nanosleep(a.poll_sleep);
    }
  };
  auto grid_sync = [&]() {
    if (CL) {
      cluster_sync_all();
    } else {
      __syncthreads();
      if (tid == 0) arrive_and_wait();
      __syncthreads();
    }
  };
  // A reduction point.  Block partial: warp shuffles, one shared-memory hop, the first warp folds the warps.
  // Grid: lane 0 stores the CTA partial into its slot and arrives; after the counter has seen every CTA the
  // FIRST WARP of every CTA sums all slots itself (fixed order, both sums interleaved) and rounds once.
  // Cluster: the partial is pushed into every CTA's inbox, then the cluster barrier.
  // red[] lands in the registers of thread 0, identical in every CTA.
  auto reduce = [&](Acc<T> e0, Acc<T> e1, bool two) {
    e0 = warp_sum(e0);
    if (two) e1 = warp_sum(e1);
    if (lane == 0) {
      wsum[0][wid] = e0;
      wsum[1][wid] = e1;
    }
    __syncthreads();
    Acc<T> b0 = zero_of<Acc<T>>(), b1 = zero_of<Acc<T>>();
    if (wid == 0) {
      b0 = lane < NW ? wsum[0][lane] : zero_of<Acc<T>>();
      b1 = lane < NW ? wsum[1][lane] : zero_of<Acc<T>>();
      b0 = warp_sum_n<NW>(b0);
      if (two) b1 = warp_sum_n<NW>(b1);
    }
    if (CL) {
      const int par = rp & 1;
      if (wid == 0) {
        b0 = shfl_idx(b0, 0);
        b1 = shfl_idx(b1, 0);
        if (lane < G) {
          const uint32_t ra = dsmem_addr(&inbox[par][blockIdx.x][0], (unsigned)lane);
          dslot_st(ra, b0);
          dslot_st(ra + 8 * AW, b1);
        }
      }
      lap(5);  // block partial
      cluster_sync_all();
      lap(6);  // waiting for the other CTAs
      if (wid == 0) {
        Acc<T> g0 = zero_of<Acc<T>>(), g1 = zero_of<Acc<T>>();
        if (lane < G) {
          sslot_ld(&inbox[par][lane][0], g0);
          sslot_ld(&inbox[par][lane][AW], g1);
        }
        g0 = warp_sum_n<kFusedCluster>(g0);
        if (two) g1 = warp_sum_n<kFusedCluster>(g1);
        red[0] = acc_round<T>(g0);
        red[1] = acc_round<T>(g1);
      }
      rp ^= 1;
    } else {
      double* const buf = a.rslots + (size_t)rp * G * 2 * AW;
      if (tid == 0) {
        slot_st(buf + (size_t)blockIdx.x * 2 * AW, b0);
        slot_st(buf + (size_t)blockIdx.x * 2 * AW + AW, b1);
        lap(5);  // block partial
        arrive_and_wait();
        lap(6);  // waiting for the other CTAs
      }
      if (wid == 0) {
        __syncwarp();
        Acc<T> g0 = zero_of<Acc<T>>(), g1 = zero_of<Acc<T>>();
#pragma unroll
        for (int k = 0; k < MAXS; ++k) {
          const int i = lane + 32 * k;
          if (i < G) {
            Acc<T> q0, q1;
            slot_ld(buf + (size_t)i * 2 * AW, q0);
            g0 = add(g0, q0);
            if (two) {
              slot_ld(buf + (size_t)i * 2 * AW + AW, q1);
              g1 = add(g1, q1);
            }
          }
        }
        g0 = warp_sum(g0);
        if (two) g1 = warp_sum(g1);
        red[0] = acc_round<T>(g0);
        red[1] = acc_round<T>(g1);  // (valid in lane 0 = thread 0, the only consumer)
      }
      rp = rp == 2 ? 0 : rp + 1;
    }
    lap(7);  // sum of the CTA partials
  };
  // The own part of the gathered vector is written (window + the rows the neighbours need in global memory, or
  // the whole vector in MODE 0 / 1): wait until what this CTA gathers from the others is there.
  auto gather_ready = [&]() {
    if (CL) {
      cluster_sync_all();  // every CTA's own part is in its window
      for (int h = tid; h < nlo + nhi; h += BLOCK) {
        const int j = h < nlo ? wlo + h : row1 + (h - nlo);
        const int owner = j / R;
        const int owlo = max(0, owner * R - a.bw_lo);
        win[j - wlo] = dsmem_ld(dsmem_addr(win + (j - owlo), (unsigned)owner), (const T*)nullptr);
      }
      __syncthreads();
    } else {
      grid_sync();
    }
  };
  // r = A x - rhs ; r0 = r ; ||r||^2   (:243-251, and the restart :305-316)
  auto residual = [&](int restart) {
    if (SV && restart) {  // x lives in shared memory: the neighbours gather it from global memory
      SPB_ROWS(l, i) st_l2(a.x + i, x_v[l]);
      grid_sync();
    }
    Acc<T> e0 = zero_of<Acc<T>>();
    const T m1 = neg(one_of<T>());
    SPB_ROWS(l, i) {
      const int2 e = extent(l, i);
      const T ri = add(fused_row(a, e.x, e.y, a.x), mul(a.rhs[i], m1));
      r_v[SPB_OWN(l, i)] = ri;
      r0_v[SPB_OWN(l, i)] = ri;
      acc_sq(e0, ri);
    }
    reduce(e0, zero_of<Acc<T>>(), false);
    if (tid == 0) bicg_s_init_body(&S, red, hist, a.cap, restart);
    __syncthreads();
  };
  // everything of an iteration after the S1 test; returns false when the solve ended (breakdown)
  auto iteration = [&](bool first) -> bool {
    {  // K1: p, y = M p (y is what the SpMV gathers)
      const T c_pv = S.c_pv, beta = S.beta, one = one_of<T>();
      SPB_ROWS(l, i) {
        const int o = SPB_OWN(l, i);
        T pi;
        if (first) {
          pi = r_v[o];
        } else {
          pi = add(mul(v_v[o], c_pv), mul(p_v[o], beta));
          pi = add(pi, mul(r_v[o], one));
        }
        p_v[o] = pi;
        const T yi = PC ? mul_diag(pi, d_v[o]) : pi;
        if (WIN) {
          win[i - wlo] = yi;
        } else {
          a.y[i] = yi;
        }
      }
      if (SV && !WIN && tid < nrows) {  // first row of the SpMV that follows the barrier
        const int2 e0 = extent(tid, row0 + tid);
        fused_prefetch_row(a, e0.x, e0.y);
      }
    }
    lap(0);  // K1
    gather_ready();
    lap(8);  // y of the neighbours / plain barrier
    {  // v = A y, <r0, v>
      Acc<T> e0 = zero_of<Acc<T>>();
      if (WIN) {
        fused_spmv_win<T, BLOCK>(a, ext_v, nrows, row0, win, wlo, mcols, mvals, mbase, mat_smem, [&](int l, int, T vi) {
          v_v[l] = vi;
          acc_prod(e0, conj_of(r0_v[l]), vi);
        });
      } else {
        SPB_ROWS(l, i) {
          const int o = SPB_OWN(l, i);
          const int2 e = extent(l, i);
          if (SV && l + BLOCK < nrows) {
            const int2 en = extent(l + BLOCK, i + BLOCK);
            fused_prefetch_row(a, en.x, en.y);
          }
          const T vi = fused_row(a, e.x, e.y, a.y);
          v_v[o] = vi;
          acc_prod(e0, conj_of(r0_v[o]), vi);
        }
      }
      lap(1);  // SpMV 1
      reduce(e0, zero_of<Acc<T>>(), false);
    }
    if (tid == 0) bicg_s2_body(&S, red, first ? 1 : 0);
    __syncthreads();
    lap(9);  // scalar step
    if (S.h.status != DS_RUNNING) return false;
    {  // K2: r -= alpha v, z = M r
      const T nalpha = S.nalpha;
      SPB_ROWS(l, i) {
        const int o = SPB_OWN(l, i);
        const T ri = add(r_v[o], mul(v_v[o], nalpha));
        r_v[o] = ri;
        const T zi = PC ? mul_diag(ri, d_v[o]) : ri;
        if (WIN) {
          win[i - wlo] = zi;
        } else {
          a.z[i] = zi;
        }
      }
      if (SV && !WIN && tid < nrows) {
        const int2 e0 = extent(tid, row0 + tid);
        fused_prefetch_row(a, e0.x, e0.y);
      }
    }
    lap(2);  // K2
    gather_ready();
    lap(8);
    {  // t = A z, <t,t>, <t,r>
      Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
      if (WIN) {
        fused_spmv_win<T, BLOCK>(a, ext_v, nrows, row0, win, wlo, mcols, mvals, mbase, mat_smem, [&](int l, int, T ti) {
          t_v[l] = ti;
          const T cy = conj_of(ti);
          acc_prod(e0, cy, ti);
          acc_prod(e1, cy, r_v[l]);
        });
      } else {
        SPB_ROWS(l, i) {
          const int o = SPB_OWN(l, i);
          const int2 e = extent(l, i);
          if (SV && l + BLOCK < nrows) {
            const int2 en = extent(l + BLOCK, i + BLOCK);
            fused_prefetch_row(a, en.x, en.y);
          }
          const T ti = fused_row(a, e.x, e.y, a.z);
          t_v[o] = ti;
          const T cy = conj_of(ti);
          acc_prod(e0, cy, ti);
          acc_prod(e1, cy, r_v[o]);
        }
      }
      lap(3);  // SpMV 2
      reduce(e0, e1, true);
    }
    if (tid == 0) bicg_s3_body(&S, red);
    __syncthreads();
    lap(9);
    {  // K3 + the partials of the next iteration's test.  y_i, z_i of the own rows are recomputed from p and r
       // (the same single operation on the same operands: the same bits) instead of re-read from memory.
      Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
      const T nalpha = S.nalpha, nw = S.nw;
      SPB_ROWS(l, i) {
        const int o = SPB_OWN(l, i);
        const T pi = p_v[o], si = r_v[o];
        const T yi = PC ? mul_diag(pi, d_v[o]) : pi;
        const T zi = PC ? mul_diag(si, d_v[o]) : si;
        T xi = add(x_v[o], mul(yi, nalpha));
        xi = add(xi, mul(zi, nw));
        x_v[o] = xi;
        const T ri = add(si, mul(t_v[o], nw));
        r_v[o] = ri;
        acc_sq(e0, ri);
        acc_prod(e1, conj_of(r0_v[o]), ri);
      }
      lap(4);  // K3
      reduce(e0, e1, true);
    }
    return true;
  };

  if (tid == 0) {
    memset(&S, 0, sizeof(S));
    S.h.status = DS_RUNNING;
    S.tol = (real_t<T>)a.tol;
  }
  if (SV) {  // row extents, 1 / diag and x of the own rows
    int2* ew = reinterpret_cast<int2*>(fused_smem + off_e);
    V* dw = reinterpret_cast<V*>(fused_smem + off_d);
    SPB_ROWS(l, i) {
      ew[l] = make_int2(__ldg(a.indptr + i), __ldg(a.indptr + i + 1));
      if (PC) dw[l] = dinv_g[i];
      x_v[l] = a.x[i];
    }
  }
  if (mat_smem) {
    const int cnt = __ldg(a.indptr + row1) - mbase;
    for (int k = tid; k < cnt; k += BLOCK) {
      mcols[k] = __ldg(a.cols + mbase + k) - wlo;
      mvals[k] = ld_ro(a.vals + mbase + k);
    }
  }
  if (CL) cluster_sync_all();  // every CTA of the cluster runs before its shared memory is addressed
  __syncthreads();
  {  // ||b||  (:225-231)
    Acc<T> e0 = zero_of<Acc<T>>();
    SPB_ROWS(l, i) acc_sq(e0, a.rhs[i]);
    reduce(e0, zero_of<Acc<T>>(), false);
    if (tid == 0) bicg_s_rhs_body(&S, red);
    __syncthreads();
  }
  if (S.h.status == DS_ZERO_RHS) {
    SPB_ROWS(l, i) x_v[SPB_OWN(l, i)] = zero_of<T>();
  } else {
    residual(0);
    if (S.h.status == DS_RUNNING && iteration(true)) {
      for (long long k = 1; k < a.max_iter; ++k) {  // :295
        if (tid == 0) bicg_s1_body(&S, red, hist, a.cap);
        __syncthreads();
        lap(9);
        if (S.h.status == DS_NEED_RESTART) residual(1);  // sets the status back to running
        if (S.h.status != DS_RUNNING) break;
        if (!iteration(false)) break;
      }
    }
  }
  if (SV) SPB_ROWS(l, i) a.x[i] = x_v[l];
  if (blockIdx.x == 0 && tid == 0) *a.st = S;
  if (a.stats && tid == 0)
    for (int k = 0; k < 10; ++k) a.stats[10 * blockIdx.x + k] = t_acc[k];
  if (CL) cluster_sync_all();  // no CTA leaves while its shared memory may still be read
#undef SPB_OWN
#undef SPB_ROWS
}

// max (row - col) and max (col - row): the reach of the gathers below / above the diagonal
__global__ void fused_band_kernel(const int* indptr, const int* cols, int n, int* out) {
  int lo = 0, hi = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    for (int k = indptr[r]; k < indptr[r + 1]; ++k) {
      const int c = cols[k];
      lo = max(lo, r - c);
      hi = max(hi, c - r);
    }
  if (lo) atomicMax(out, lo);  // integer atomics: order-independent
  if (hi) atomicMax(out + 1, hi);
}

// ---------------------------------------------------------------- host driver
template <typename T>
struct BicgStab : spb_solver {
  DevBuf ws;        // 7 n T  (src/bicg_stab.rs:28)
  DevBuf partials;  // Acc<T> [2 * max grid]
  DevBuf red;       // scal2 [2]
  DevBuf state;     // BicgState<T>
  DevBuf hist_d;
  DevBuf fused_slots, fused_bar, fused_stats, fused_band;  // single-kernel path
  int band_lo = -1, band_hi = -1;  // reach of the gathers below / above the diagonal (computed once, lazily)

  // Single-kernel solve: one GPU, Jacobi or no preconditioner, matrix + vectors resident in L2.
  bool fused_eligible(const CsrMat<T>* Am, PcMode pcm) const {
    const char* e = getenv("SPB_FUSED");
    if (e && *e == '0') return false;
    if (ctx->dist || Am->ip64 || Am->n_halo > 0 || pcm == PCM_GENERIC || size <= 0) return false;
    if (e && *e == '1') return true;
    const char* mb = getenv("SPB_FUSED_MAX_MB");
    const double limit = (mb && *mb ? atof(mb) : 48.0) * 1048576.0;
    const double bytes = (double)Am->nnz * (sizeof(T) + 4) + 10.0 * (double)size * sizeof(T);
    return bytes <= limit;
  }
  template <typename V, bool PC>
  void launch_fused(const FusedArgs<T>& fa0, int64_t n);

  BicgStab(spb_op* A_, int64_t size_) {
    A = A_;
    ctx = A_->ctx;
    kind = 0;
    dtype = ScalarTraits<T>::dtype;
    size = size_;
    ws.alloc(sizeof(T) * 7 * (size_t)std::max<int64_t>(size, 1));
    SPB_CUDA(cudaMemsetAsync(ws.p, 0, ws.bytes, ctx->stream));
    partials.alloc(sizeof(Acc<T>) * 2 * (size_t)(vec_max_grid(ctx) + 1));
    red.alloc(sizeof(scal2) * 2);
    state.alloc(sizeof(BicgState<T>));
  }

  int solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double tol, int64_t* iters,
                double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) override;
};

template <typename T>
int BicgStab<T>::solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double tol,
                           int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                           int64_t* hist_len) {
  Ctx* c = ctx;
  const int64_t n = size;
  const T* rhs = (const T*)d_rhs;
  T* x = (T*)d_x;
  if (A->kind != OP_CSR) SPB_FAIL(SPB_INVALID_ARG, "BiCGStab operator must be a CSR matrix");
  auto* Am = static_cast<CsrMat<T>*>(A);
  if (Am->n_local != n) {
    set_last_error("Input vec dimension doesn't match the matrix size");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  if (M && M->n_local != n) SPB_FAIL(SPB_DIM_MISMATCH, "preconditioner dimension mismatch");
  const PcMode pcm = pc_mode_of<T>(M);
  T* w0 = bufptr<T>(ws);
  T *r = w0, *r0 = w0 + n, *p = w0 + 2 * n, *y = w0 + 3 * n, *v = w0 + 4 * n, *t = w0 + 5 * n, *z = w0 + 6 * n;
  if (pcm == PCM_NONE) {  // solve(): y is p, z is r (src/bicg_stab.rs:65-71, 116)
    y = p;
    z = r;
  }
  const void* dinv = (pcm == PCM_JACOBI || pcm == PCM_JACOBI_REAL) ? static_cast<DiagOp<T>*>(M)->dinv.p : nullptr;
  auto* st = bufptr<BicgState<T>>(state);
  scal2* redp = bufptr<scal2>(red);
  Acc<T>* parts = bufptr<Acc<T>>(partials);
  const int grid = vec_grid(c, n);
  const int grid_r = std::min(vec_grid_resident(c, n, bicg_k3<T>), vec_grid_resident(c, n, bicg_k_init<T>));  // kernels with partial sums
  const long long cap = hist ? std::min<int64_t>(hist_cap, max_iter + 1) : 0;
  double* hd = nullptr;
  if (cap > 0) {
    hist_d.ensure(sizeof(double) * cap);
    hd = bufptr<double>(hist_d);
  }

  BicgState<T> init;
  memset(&init, 0, sizeof(init));
  init.h.status = DS_RUNNING;
  init.tol = (real_t<T>)tol;
  SPB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));

  if (fused_eligible(Am, pcm)) {
    // the whole solve in one cooperative kernel (see bicg_fused_kernel)
    // (y and z are always separate buffers here: they are what the other CTAs gather)
    FusedArgs<T> fa{bufptr<int>(Am->indptr), bufptr<int>(Am->cols), bufptr<T>(Am->vals), (int)n, rhs, x, r, r0, p, w0 + 3 * n, v, t, w0 + 6 * n,
                    dinv, st, nullptr, nullptr, hd, cap, max_iter, tol, nullptr, 0, 0, 0, 0, 0, 0, 0};
    if (const char* ps = getenv("SPB_FUSED_POLL_NS")) fa.poll_sleep = (unsigned)atoi(ps);
    const bool want_stats = getenv("SPB_FUSED_STATS") != nullptr;
    if (want_stats) {
      fused_stats.ensure(sizeof(long long) * 10 * kFusedMaxGrid);
      SPB_CUDA(cudaMemsetAsync(fused_stats.p, 0, sizeof(long long) * 10 * kFusedMaxGrid, c->stream));
      fa.stats = bufptr<long long>(fused_stats);
    }
    c->gate = nullptr;
    if (pcm == PCM_JACOBI)
      launch_fused<T, true>(fa, n);
    else if (pcm == PCM_JACOBI_REAL)
      launch_fused<real_t<T>, true>(fa, n);
    else
      launch_fused<T, false>(fa, n);
    BicgState<T> fin;
    SPB_CUDA(cudaMemcpyAsync(&fin, st, sizeof(fin), cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    if (want_stats) {
      std::vector<long long> hs(10 * kFusedMaxGrid);
      SPB_CUDA(cudaMemcpy(hs.data(), fused_stats.p, sizeof(long long) * hs.size(), cudaMemcpyDeviceToHost));
      static const char* nm[10] = {"K1", "SpMV1", "K2", "SpMV2", "K3", "block partial", "reduction wait", "partial sum", "gather hand-off", "scalar step"};
      const double its = (double)std::max<long long>(fin.h.its, 1);
      int ng = 0;
      for (int b = 0; b < kFusedMaxGrid; ++b)
        if (hs[10 * b + 5] > 0) ng = b + 1;
      fprintf(stderr, "[bicg_fused] clocks per iteration, min / median / max over %d CTAs:", ng);
      for (int k = 0; k < 10; ++k) {
        std::vector<double> v;
        for (int b = 0; b < ng; ++b) v.push_back((double)hs[10 * b + k] / its);
        std::sort(v.begin(), v.end());
        if (!v.empty()) fprintf(stderr, " %s=%.0f/%.0f/%.0f", nm[k], v.front(), v[v.size() / 2], v.back());
      }
      fprintf(stderr, "\n");
    }
    int rcf;
    if (fin.h.status == DS_OK || fin.h.status == DS_ZERO_RHS) {
      *iters = fin.h.res_iters;
      *resid = fin.h.res_resid;
      rcf = SPB_OK;
    } else if (fin.h.status == DS_BREAKDOWN) {
      *iters = fin.h.res_iters;
      rcf = SPB_BREAKDOWN;
    } else {
      *iters = max_iter;
      rcf = SPB_INSUFFICIENT_ITER;
    }
    const int64_t hl = fin.h.status == DS_ZERO_RHS ? 0 : fin.h.hist_len;
    if (hist_len) *hist_len = hl;
    if (hd && hl > 0)
      SPB_CUDA(cudaMemcpy(hist, hd, sizeof(double) * std::min<int64_t>(hl, cap), cudaMemcpyDeviceToHost));
    return rcf;
  }

  auto scalar = [&](auto kernel, auto... args) {
    LaunchScope ls(c, FAM_SCALAR);
    kernel<<<1, 1, 0, c->stream>>>(args...);
    check_launch("bicg scalar kernel");
  };
  auto reduce_vec = [&]() {  // per-block partials of a vector kernel -> red (all ranks)
    finalize_allreduce<T>(c, parts, grid_r, redp);
  };
  auto reduce_spmv = [&](auto tail) {  // epilogue partials of the last SpMV -> Am->red (all ranks), then the scalar step
    finalize_reduce_tail<T>(c, bufptr<Acc<T>>(Am->partials), Am->last_partial_blocks, bufptr<scal2>(Am->red), true, tail);
  };
  auto k_init = [&](int restart) {
    LaunchScope ls(c, FAM_VEC);
    bicg_k_init<T><<<grid_r, kVecThreads, 0, c->stream>>>(st, restart, n, rhs, r, r0, parts);
    check_launch("bicg_k_init");
  };
  auto k1 = [&](bool first) {
    {
      LaunchScope ls(c, FAM_VEC);
      if (pcm == PCM_JACOBI) {
        if (first) bicg_k1<T, T, true, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const T*)dinv);
        else bicg_k1<T, T, false, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const T*)dinv);
      } else if (pcm == PCM_JACOBI_REAL) {
        if (first) bicg_k1<T, real_t<T>, true, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const real_t<T>*)dinv);
        else bicg_k1<T, real_t<T>, false, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const real_t<T>*)dinv);
      } else {
        if (first) bicg_k1<T, T, true, false><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const T*)nullptr);
        else bicg_k1<T, T, false, false><<<grid, kVecThreads, 0, c->stream>>>(st, n, v, p, r, y, (const T*)nullptr);
      }
      check_launch("bicg_k1");
    }
    if (pcm == PCM_GENERIC) op_apply<T>(M, p, y);
  };
  auto k2 = [&]() {
    {
      LaunchScope ls(c, FAM_VEC);
      if (pcm == PCM_JACOBI) bicg_k2<T, T, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, r, v, z, (const T*)dinv);
      else if (pcm == PCM_JACOBI_REAL) bicg_k2<T, real_t<T>, true><<<grid, kVecThreads, 0, c->stream>>>(st, n, r, v, z, (const real_t<T>*)dinv);
      else bicg_k2<T, T, false><<<grid, kVecThreads, 0, c->stream>>>(st, n, r, v, z, (const T*)nullptr);
      check_launch("bicg_k2");
    }
    if (pcm == PCM_GENERIC) op_apply<T>(M, r, z);
  };
  auto tail = [&](bool first) {  // everything of an iteration after the S1 test
    k1(first);
    Am->mul(y, v, EPI_DOT_WY, r0, false);
    reduce_spmv(BicgS2Tail<T>{st, bufptr<scal2>(Am->red), first ? 1 : 0});  // + alpha = rho / <r0, v>
    k2();
    Am->mul(z, t, EPI_TT_TR, r, false);
    reduce_spmv(BicgS3Tail<T>{st, bufptr<scal2>(Am->red)});                 // + w = <t, r> / <t, t>
    {
      LaunchScope ls(c, FAM_VEC);
      bicg_k3<T><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, x, y, z, r, t, r0, parts);
      check_launch("bicg_k3");
    }
    reduce_vec();
  };

  int rc = SPB_OK;
  Poller poller(c);
  StateHead hd_host;
  c->gate = nullptr;
  try {
    // ||b||                                                           (:225-231)
    vec_reduce<T>(c, 2, n, rhs, rhs, parts, redp, true);
    scalar(bicg_s_rhs<T>, st, redp);
    poller.post(st);
    poller.drain(&hd_host);
    if (hd_host.status == DS_ZERO_RHS) {
      SPB_CUDA(cudaMemsetAsync(x, 0, sizeof(T) * n, c->stream));  // x := 0
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      *iters = 0;
      *resid = hd_host.res_resid;
      if (hist_len) *hist_len = 0;
      return SPB_OK;
    }
    c->gate = &st->h.status;
    Am->mul(x, r, EPI_NONE, nullptr, false);  // r = A x            (:244)
    k_init(0);                                // r -= rhs, r0 = r   (:246-251)
    reduce_vec();
    scalar(bicg_s_init<T>, st, redp, hd, cap, 0);
    tail(true);  // unrolled first iteration                         (:260-293)

    const int64_t target = max_iter - 1;  // loop iterations 1 .. max_iter-1  (:295)
    int64_t launched = 0;                 // S1 launches that took effect
    bool done = false;
    auto launch_iter = [&]() {
      scalar(bicg_s1<T>, st, redp, hd, cap);
      tail(false);
      ++launched;
    };
    auto handle = [&](const StateHead& h) {
      if (h.status == DS_RUNNING) return;
      if (h.status == DS_NEED_RESTART) {
        // rho restart (:304-318).  Everything queued after the pausing S1 was a no-op; drain
        // it, redo r = A x - rhs, r0 = r, rho, beta on the device and resume iteration h.its.
        StateHead last;
        poller.drain(&last);
        c->gate_value = DS_NEED_RESTART;
        Am->mul(x, r, EPI_NONE, nullptr, false);
        k_init(1);
        reduce_vec();
        scalar(bicg_s_init<T>, st, redp, hd, cap, 1);
        c->gate_value = DS_RUNNING;
        launched = h.its;
        tail(false);
        return;
      }
      done = true;
    };
    while (!done) {
      StateHead h;
      if (launched < target) {
        const int64_t chunk = std::min<int64_t>(poll, target - launched);
        for (int64_t i = 0; i < chunk; ++i) launch_iter();
        poller.post(st);
        if (poller.wait_oldest(&h)) handle(h);  // lags one chunk behind: the queue never drains
      } else {
        if (!poller.drain(&h)) {
          poller.post(st);
          poller.drain(&h);
        }
        handle(h);
        if (!done && h.status == DS_RUNNING && launched >= target) break;  // exhausted (:365)
      }
    }
    c->gate = nullptr;
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    // final, authoritative read of the state
    poller.post(st);
    poller.drain(&hd_host);
    if (hd_host.status == DS_OK) {
      *iters = hd_host.res_iters;
      *resid = hd_host.res_resid;
      rc = SPB_OK;
    } else if (hd_host.status == DS_BREAKDOWN) {
      *iters = hd_host.res_iters;
      rc = SPB_BREAKDOWN;
    } else {
      *iters = max_iter;
      rc = SPB_INSUFFICIENT_ITER;
    }
    const int64_t hl = hd_host.hist_len;
    if (hist_len) *hist_len = hl;
    if (hd && hl > 0)
      SPB_CUDA(cudaMemcpy(hist, hd, sizeof(double) * std::min<int64_t>(hl, cap), cudaMemcpyDeviceToHost));
  } catch (...) {
    c->gate = nullptr;
    throw;
  }
  return rc;
}

template <typename T>
template <typename V, bool PC>
void BicgStab<T>::launch_fused(const FusedArgs<T>& fa0, int64_t n) {
  Ctx* c = ctx;
  FusedArgs<T> fa = fa0;
  const char* be = getenv("SPB_FUSED_BLOCK");
  // 512 threads per CTA; 256 for the smallest systems (<= 8192 rows: 64^2 7.6 vs 8.1 us per iteration in cluster mode)
  const int block = be && *be ? atoi(be) : (n <= 8192 ? 256 : 512);
  const char* se = getenv("SPB_FUSED_SMEM");
  const bool allow_smem = !(se && *se == '0');
  const bool allow_win = allow_smem;  // (the cluster mode keeps the own-row vectors in shared memory too)
  int smem_cap = 0;
  SPB_CUDA(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  if (allow_win && band_lo < 0) {  // bandwidth of the matrix: decides whether the gathered vector fits a shared-memory window
    fused_band.ensure(sizeof(int) * 2);
    SPB_CUDA(cudaMemsetAsync(fused_band.p, 0, sizeof(int) * 2, c->stream));
    {
      LaunchScope ls(c, FAM_SCALAR);
      fused_band_kernel<<<(int)std::min<int64_t>(ceil_div(n, 256), 1024), 256, 0, c->stream>>>(fa.indptr, fa.cols, (int)n, bufptr<int>(fused_band));
      check_launch("fused_band_kernel");
    }
    int hb[2] = {0, 0};
    SPB_CUDA(cudaMemcpyAsync(hb, fused_band.p, sizeof(hb), cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    band_lo = hb[0];
    band_hi = hb[1];
  }
  const char* ce = getenv("SPB_FUSED_CLUSTER");
  const bool allow_cluster = allow_win && !(ce && *ce == '0');
  auto run = [&](auto kern_cl, auto kern_sv, auto kern_gl, int BLOCK) {
    constexpr int AW = (int)sizeof(Acc<T>) / 8;
    const size_t per_slot = (size_t)6 * sizeof(T) + (PC ? sizeof(V) : 0) + sizeof(int2);
    fa.bw_lo = std::max(band_lo, 0);
    fa.bw_hi = std::max(band_hi, 0);
    fused_bar.ensure(sizeof(unsigned long long) * 2);
    fa.bar = bufptr<unsigned long long>(fused_bar);
    SPB_CUDA(cudaMemsetAsync(fused_bar.p, 0, sizeof(unsigned long long) * 2, c->stream));
    // ---- MODE 2: the whole system inside one thread-block cluster (<= 4 rows per thread)
    if (allow_cluster && band_lo >= 0) {
      const int gc = (int)std::max<int64_t>(1, std::min<int64_t>(kFusedCluster, ceil_div(n, BLOCK)));
      const int R = (int)ceil_div(n, (int64_t)gc);
      const int spt = (int)ceil_div((int64_t)R, (int64_t)BLOCK);
      const int slots = (R + 31) & ~31;
      const size_t smem_sv = (per_slot * (size_t)slots + 15) / 16 * 16;
      const int64_t win_elems = (int64_t)R + fa.bw_lo + fa.bw_hi;
      const size_t smem_win = (smem_sv + (size_t)win_elems * sizeof(T) + 31) / 16 * 16;
      // matrix slice: mean entries per CTA + 25 % (a CTA whose slice is larger reads its matrix from global memory)
      const int64_t nnz_all = static_cast<CsrMat<T>*>(A)->nnz;
      int64_t mat_cap = (nnz_all / gc) + (nnz_all / gc) / 4 + 64;
      size_t smem_all = smem_win + (size_t)mat_cap * (sizeof(int) + sizeof(T)) + 48;
      if (smem_all + 4096 > (size_t)smem_cap) {
        mat_cap = 0;
        smem_all = smem_win + 48;
      }
      const bool force = ce && *ce == '1';
      if ((spt <= 4 || force) && smem_all + 4096 <= (size_t)smem_cap) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(gc);
        cfg.blockDim = dim3(BLOCK);
        cfg.dynamicSmemBytes = smem_all;
        cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = gc;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        const bool ok = cudaFuncSetAttribute(kern_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all) == cudaSuccess &&
                        cudaFuncSetAttribute(kern_cl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaOccupancyMaxActiveClusters(&nclusters, kern_cl, &cfg) == cudaSuccess && nclusters >= 1;
        if (ok) {
          fa.slots = slots;
          fa.rows_per_cta = R;
          fa.win_elems = (int)win_elems;
          fa.mat_cap = (int)mat_cap;
          LaunchScope ls(c, FAM_VEC);
          SPB_CUDA(cudaLaunchKernelEx(&cfg, kern_cl, fa));
          return;
        }
        cudaGetLastError();  // the cluster shape is not available on this device: use the grid modes
      }
    }
    // ---- MODE 0 / 1: cooperative grid, one CTA per SM (the shared-memory variant needs most of an SM's shared memory anyway)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(c->sm_count, kFusedMaxGrid), ceil_div(n, BLOCK)));
    const int R = (int)ceil_div(n, (int64_t)grid);
    const int slots = (R + 31) & ~31;
    const size_t smem_sv = (per_slot * (size_t)slots + 15) / 16 * 16 + 64;
    const bool sv = allow_smem && smem_sv + 4096 <= (size_t)smem_cap;
    fa.slots = slots;
    fa.rows_per_cta = R;
    fa.win_elems = 0;
    fa.mat_cap = 0;
    fused_slots.ensure(sizeof(double) * 3 * 2 * AW * (size_t)grid);
    fa.rslots = bufptr<double>(fused_slots);
    LaunchScope ls(c, FAM_VEC);
    void* args[] = {(void*)&fa};
    if (sv) {
      SPB_CUDA(cudaFuncSetAttribute(kern_sv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sv));
      SPB_CUDA(cudaLaunchCooperativeKernel((void*)kern_sv, dim3(grid), dim3(BLOCK), args, smem_sv, c->stream));
    } else {
      SPB_CUDA(cudaLaunchCooperativeKernel((void*)kern_gl, dim3(grid), dim3(BLOCK), args, 0, c->stream));
    }
  };
  if (block >= 512)
    run(bicg_fused_kernel<T, V, PC, 512, 2>, bicg_fused_kernel<T, V, PC, 512, 1>, bicg_fused_kernel<T, V, PC, 512, 0>, 512);
  else
    run(bicg_fused_kernel<T, V, PC, 256, 2>, bicg_fused_kernel<T, V, PC, 256, 1>, bicg_fused_kernel<T, V, PC, 256, 0>, 256);
}

spb_solver* make_bicgstab(spb_op* A, int64_t size) {
  switch (A->dtype) {
    case SPB_F64: return new BicgStab<double>(A, size);
    case SPB_C128: return new BicgStab<cplx>(A, size);
    case SPB_F32: return new BicgStab<float>(A, size);
    default: return new BicgStab<cplxf>(A, size);
  }
}

}  // namespace spb
