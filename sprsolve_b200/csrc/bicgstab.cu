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
// cooperative kernel: one CTA per SM, a thread owns the same rows in every phase, phases are separated
// by a grid barrier (one atomic + one spin per CTA) only where a vector written by other CTAs is
// gathered (before each SpMV) or a reduction completes.  Every CTA sums the per-CTA double-double
// partials of a reduction point itself (same fixed order, one rounding), so the Krylov scalars are
// replicated in shared memory and no second barrier or broadcast is needed.
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
  Acc<T>* parts;            // [3][2 * gridDim.x]
  unsigned long long* bar;  // grid barrier counter, zeroed before the launch
  double* hist;
  long long cap, max_iter;
  double tol;
  long long* stats;  // optional (SPB_FUSED_STATS=1): clocks of CTA 0 per phase, see the lap() calls
  int slots_per_thread;  // ceil(n / threads of the grid): shared-memory slots a thread needs per vector
  unsigned poll_sleep;   // ns between two polls of the grid barrier (0: spin)
};

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// vectors other CTAs write during the kernel are read at L2 (the L1 of an SM is not coherent)
__device__ __forceinline__ double ld_l2(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_l2(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}
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
__device__ __forceinline__ AccR ld_l2(const AccR* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return AccR{v.x, v.y};
}
__device__ __forceinline__ AccC ld_l2(const AccC* p) {
  const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
  return AccC{a.x, a.y, b.x, b.y};
}
template <typename T>
__device__ __forceinline__ scal2 acc_round(const AccR& a) {
  return scal2{round_to_real<T>(a.hi + a.lo), 0.0};
}
template <typename T>
__device__ __forceinline__ scal2 acc_round(const AccC& a) {
  return scal2{round_to_real<T>(a.rh + a.rl), round_to_real<T>(a.ih + a.il)};
}

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

// SV: the vectors only their owning thread touches (r, r0, p, v, t, 1/diag) live in SHARED MEMORY for the
// whole solve -- a thread owns rows gtid + k * nth, slot k * BLOCK + tid -- so the vector-update phases
// never wait for L2 (~1000 clocks per dependent access under this load); only what other CTAs gather
// (y, z) and x go to global memory.  Falls back to global vectors (SV = false) when the slots do not fit.
template <typename T, typename V, bool PC, int BLOCK, bool SV>
__global__ void __launch_bounds__(BLOCK, 1) bicg_fused_kernel(const FusedArgs<T> a) {
  constexpr int NW = BLOCK / 32;
  extern __shared__ __align__(16) unsigned char fused_smem[];
  __shared__ BicgState<T> S;
  __shared__ Acc<T> wsum[2][NW];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gtid = blockIdx.x * BLOCK + tid, nth = gridDim.x * BLOCK, n = a.n;
  const V* dinv_g = static_cast<const V*>(a.dinv);
  double* hist = blockIdx.x == 0 ? a.hist : nullptr;  // one writer
  unsigned long long target = 0;
  int rp = 0;  // reduction points cycle through three partial buffers (a buffer is re-written two barriers later)
  scal2 red[2];
  // own-row vectors: shared memory (slot = k * BLOCK + tid) or global (index = row)
  const int slots = a.slots_per_thread * BLOCK;
  T* const sm = reinterpret_cast<T*>(fused_smem);
  T* const r_v = SV ? sm : a.r;
  T* const r0_v = SV ? sm + slots : a.r0;
  T* const p_v = SV ? sm + 2 * slots : a.p;
  T* const v_v = SV ? sm + 3 * slots : a.v;
  T* const t_v = SV ? sm + 4 * slots : a.t;
  const V* const d_v = SV ? reinterpret_cast<const V*>(sm + 5 * slots) : dinv_g;
  // row extents of the own rows (shared memory: one dependent L2 round trip less per SpMV row)
  const int2* const ext_v = reinterpret_cast<const int2*>(fused_smem + (size_t)slots * (5 * sizeof(T) + (PC ? sizeof(V) : 0)));
  auto extent = [&](int k, int i) -> int2 {
    if (SV) return ext_v[k * BLOCK + tid];
    return make_int2(__ldg(a.indptr + i), __ldg(a.indptr + i + 1));
  };
#define SPB_OWN(k, i) (SV ? (k) * BLOCK + tid : (i))
#define SPB_ROWS(k, i) for (int k = 0, i = gtid; i < n; ++k, i += nth)
  long long t_last = a.stats ? clock64() : 0;
  long long t_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  auto lap = [&](int k) {  // diagnostics: phase clocks of CTA 0 (accumulated in registers, stored at the end)
    if (a.stats && blockIdx.x == 0 && tid == 0) {
      const long long now = clock64();
      t_acc[k] += now - t_last;
      t_last = now;
    }
  };

  // Grid barrier: a release-add without return value (the arriving thread does not wait for the round
  // trip) and an acquire spin by one thread per CTA; __syncthreads on both sides extends the ordering to
  // the whole CTA.
  auto grid_sync = [&]() {
    __syncthreads();
    if (tid == 0) {
      target += gridDim.x;
      red_release_gpu_add(a.bar, 1ULL);
      while (ld_acquire_gpu_u64(a.bar) < target) {
        if (a.poll_sleep) __nanosleep(a.poll_sleep);
      }
    }
    __syncthreads();
  };
  // A reduction point.  Block partial: warp shuffles, one shared-memory hop, the first warp folds the
  // warps; after the barrier the FIRST WARP of every CTA sums all CTA partials itself (fixed order, both
  // slots interleaved) and rounds once: red[] lands in the registers of thread 0, identical in every CTA.
  auto reduce = [&](Acc<T> e0, Acc<T> e1, bool two) {
    e0 = warp_sum(e0);
    if (two) e1 = warp_sum(e1);
    if (lane == 0) {
      wsum[0][wid] = e0;
      wsum[1][wid] = e1;
    }
    __syncthreads();
    Acc<T>* pb = a.parts + (size_t)rp * 2 * gridDim.x;
    if (wid == 0) {
      Acc<T> b0 = lane < NW ? wsum[0][lane] : zero_of<Acc<T>>();
      Acc<T> b1 = lane < NW ? wsum[1][lane] : zero_of<Acc<T>>();
      b0 = warp_sum(b0);
      if (two) b1 = warp_sum(b1);
      if (lane == 0) {
        pb[2 * blockIdx.x] = b0;
        pb[2 * blockIdx.x + 1] = b1;
      }
    }
    lap(5);  // block partial
    grid_sync();
    lap(6);  // barrier of a reduction point
    if (wid == 0) {
      Acc<T> g0 = zero_of<Acc<T>>(), g1 = zero_of<Acc<T>>();
      for (int i = lane; i < (int)gridDim.x; i += 32) {
        g0 = add(g0, ld_l2(pb + 2 * i));
        if (two) g1 = add(g1, ld_l2(pb + 2 * i + 1));
      }
      g0 = warp_sum(g0);
      if (two) g1 = warp_sum(g1);
      red[0] = acc_round<T>(g0);
      red[1] = acc_round<T>(g1);  // (valid in lane 0 = thread 0, the only consumer)
    }
    rp = rp == 2 ? 0 : rp + 1;
    lap(7);  // sum of the CTA partials
  };
  // r = A x - rhs ; r0 = r ; ||r||^2   (:243-251, and the restart :305-316)
  auto residual = [&](int restart) {
    Acc<T> e0 = zero_of<Acc<T>>();
    const T m1 = neg(one_of<T>());
    SPB_ROWS(k, i) {
      const int2 e = extent(k, i);
      if (SV && i + nth < n) {
        const int2 en = extent(k + 1, i + nth);
        fused_prefetch_row(a, en.x, en.y);
      }
      const T ri = add(fused_row(a, e.x, e.y, a.x), mul(a.rhs[i], m1));
      r_v[SPB_OWN(k, i)] = ri;
      r0_v[SPB_OWN(k, i)] = ri;
      acc_sq(e0, ri);
    }
    reduce(e0, zero_of<Acc<T>>(), false);
    if (tid == 0) bicg_s_init_body(&S, red, hist, a.cap, restart);
    __syncthreads();
  };
  // everything of an iteration after the S1 test; returns false when the solve ended (breakdown)
  auto iteration = [&](bool first) -> bool {
    {  // K1: p, y = M p (y is what the other CTAs gather)
      const T c_pv = S.c_pv, beta = S.beta, one = one_of<T>();
      SPB_ROWS(k, i) {
        const int o = SPB_OWN(k, i);
        T pi;
        if (first) {
          pi = r_v[o];
        } else {
          pi = add(mul(v_v[o], c_pv), mul(p_v[o], beta));
          pi = add(pi, mul(r_v[o], one));
        }
        p_v[o] = pi;
        a.y[i] = PC ? mul_diag(pi, d_v[o]) : pi;
      }
      if (SV && gtid < n) {  // first row of the SpMV that follows the barrier
        const int2 e0 = extent(0, gtid);
        fused_prefetch_row(a, e0.x, e0.y);
      }
    }
    lap(0);  // K1
    grid_sync();  // y complete
    lap(8);  // plain barrier
    {  // v = A y, <r0, v>
      Acc<T> e0 = zero_of<Acc<T>>();
      SPB_ROWS(k, i) {
        const int o = SPB_OWN(k, i);
        const int2 e = extent(k, i);
        if (SV && i + nth < n) {
          const int2 en = extent(k + 1, i + nth);
          fused_prefetch_row(a, en.x, en.y);
        }
        const T vi = fused_row(a, e.x, e.y, a.y);
        v_v[o] = vi;
        acc_prod(e0, conj_of(r0_v[o]), vi);
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
      SPB_ROWS(k, i) {
        const int o = SPB_OWN(k, i);
        const T ri = add(r_v[o], mul(v_v[o], nalpha));
        r_v[o] = ri;
        a.z[i] = PC ? mul_diag(ri, d_v[o]) : ri;
      }
      if (SV && gtid < n) {
        const int2 e0 = extent(0, gtid);
        fused_prefetch_row(a, e0.x, e0.y);
      }
    }
    lap(2);  // K2
    grid_sync();  // z complete
    lap(8);
    {  // t = A z, <t,t>, <t,r>
      Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
      SPB_ROWS(k, i) {
        const int o = SPB_OWN(k, i);
        const int2 e = extent(k, i);
        if (SV && i + nth < n) {
          const int2 en = extent(k + 1, i + nth);
          fused_prefetch_row(a, en.x, en.y);
        }
        const T ti = fused_row(a, e.x, e.y, a.z);
        t_v[o] = ti;
        const T cy = conj_of(ti);
        acc_prod(e0, cy, ti);
        acc_prod(e1, cy, r_v[o]);
      }
      lap(3);  // SpMV 2
      reduce(e0, e1, true);
    }
    if (tid == 0) bicg_s3_body(&S, red);
    __syncthreads();
    lap(9);
    {  // K3 + the partials of the next iteration's test.  y_i, z_i of the own rows are recomputed from p and r
       // (the same single operation on the same operands: the same bits) instead of re-read from global.
      Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
      const T nalpha = S.nalpha, nw = S.nw;
      SPB_ROWS(k, i) {
        const int o = SPB_OWN(k, i);
        const T pi = p_v[o], si = r_v[o];
        const T yi = PC ? mul_diag(pi, d_v[o]) : pi;
        const T zi = PC ? mul_diag(si, d_v[o]) : si;
        T xi = add(a.x[i], mul(yi, nalpha));
        xi = add(xi, mul(zi, nw));
        a.x[i] = xi;
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
  if (SV) {  // row extents and 1 / diag of the own rows
    int2* ew = reinterpret_cast<int2*>(fused_smem + (size_t)slots * (5 * sizeof(T) + (PC ? sizeof(V) : 0)));
    V* dw = reinterpret_cast<V*>(sm + 5 * slots);
    SPB_ROWS(k, i) {
      ew[k * BLOCK + tid] = make_int2(__ldg(a.indptr + i), __ldg(a.indptr + i + 1));
      if (PC) dw[k * BLOCK + tid] = dinv_g[i];
    }
  }
  __syncthreads();
  {  // ||b||  (:225-231)
    Acc<T> e0 = zero_of<Acc<T>>();
    for (int i = gtid; i < n; i += nth) acc_sq(e0, a.rhs[i]);
    reduce(e0, zero_of<Acc<T>>(), false);
    if (tid == 0) bicg_s_rhs_body(&S, red);
    __syncthreads();
  }
  if (S.h.status == DS_ZERO_RHS) {
    for (int i = gtid; i < n; i += nth) a.x[i] = zero_of<T>();
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
  if (blockIdx.x == 0 && tid == 0) {
    *a.st = S;
    if (a.stats)
      for (int k = 0; k < 10; ++k) a.stats[k] = t_acc[k];
  }
#undef SPB_OWN
#undef SPB_ROWS
}

// ---------------------------------------------------------------- host driver
template <typename T>
struct BicgStab : spb_solver {
  DevBuf ws;        // 7 n T  (src/bicg_stab.rs:28)
  DevBuf partials;  // Acc<T> [2 * max grid]
  DevBuf red;       // scal2 [2]
  DevBuf state;     // BicgState<T>
  DevBuf hist_d;
  DevBuf fused_parts, fused_bar, fused_stats;  // single-kernel path

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
                    dinv, st, nullptr, nullptr, hd, cap, max_iter, tol, nullptr, 0, 0};
    if (const char* ps = getenv("SPB_FUSED_POLL_NS")) fa.poll_sleep = (unsigned)atoi(ps);
    const bool want_stats = getenv("SPB_FUSED_STATS") != nullptr;
    if (want_stats) {
      fused_stats.ensure(sizeof(long long) * 16);
      SPB_CUDA(cudaMemsetAsync(fused_stats.p, 0, sizeof(long long) * 16, c->stream));
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
      long long hs[16];
      SPB_CUDA(cudaMemcpy(hs, fused_stats.p, sizeof(hs), cudaMemcpyDeviceToHost));
      static const char* nm[10] = {"K1", "SpMV1", "K2", "SpMV2", "K3", "block partial", "reduction barrier", "partial sum", "plain barrier", "scalar step"};
      const double its = (double)std::max<long long>(fin.h.its, 1);
      fprintf(stderr, "[bicg_fused] clocks per iteration (CTA 0):");
      for (int k = 0; k < 10; ++k) fprintf(stderr, " %s=%.0f", nm[k], (double)hs[k] / its);
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
    finalize_allreduce<T>(c, parts, grid, redp);
  };
  auto reduce_spmv = [&](auto tail) {  // epilogue partials of the last SpMV -> Am->red (all ranks), then the scalar step
    finalize_reduce_tail<T>(c, bufptr<Acc<T>>(Am->partials), Am->last_partial_blocks, bufptr<scal2>(Am->red), true, tail);
  };
  auto k_init = [&](int restart) {
    LaunchScope ls(c, FAM_VEC);
    bicg_k_init<T><<<grid, kVecThreads, 0, c->stream>>>(st, restart, n, rhs, r, r0, parts);
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
      bicg_k3<T><<<grid, kVecThreads, 0, c->stream>>>(st, n, x, y, z, r, t, r0, parts);
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
  const int block = be && *be ? atoi(be) : 512;
  const char* se = getenv("SPB_FUSED_SMEM");
  const bool allow_smem = !(se && *se == '0');
  int smem_cap = 0;
  SPB_CUDA(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  auto run = [&](auto kern_sv, auto kern_gl, int BLOCK) {
    // one CTA per SM (the shared-memory variant needs most of an SM's shared memory anyway)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)c->sm_count, ceil_div(n, BLOCK)));
    const int spt = (int)ceil_div(n, (int64_t)grid * BLOCK);
    const size_t smem = ((size_t)5 * sizeof(T) + (PC ? sizeof(V) : 0) + sizeof(int2)) * (size_t)spt * BLOCK + 16;
    const bool sv = allow_smem && smem + 4096 <= (size_t)smem_cap;
    fa.slots_per_thread = spt;
    fused_parts.ensure(sizeof(Acc<T>) * 6 * (size_t)grid);
    fused_bar.ensure(sizeof(unsigned long long) * 2);
    fa.parts = bufptr<Acc<T>>(fused_parts);
    fa.bar = bufptr<unsigned long long>(fused_bar);
    SPB_CUDA(cudaMemsetAsync(fused_bar.p, 0, sizeof(unsigned long long) * 2, c->stream));
    LaunchScope ls(c, FAM_VEC);
    void* args[] = {(void*)&fa};
    if (sv) {
      SPB_CUDA(cudaFuncSetAttribute(kern_sv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SPB_CUDA(cudaLaunchCooperativeKernel((void*)kern_sv, dim3(grid), dim3(BLOCK), args, smem, c->stream));
    } else {
      SPB_CUDA(cudaLaunchCooperativeKernel((void*)kern_gl, dim3(grid), dim3(BLOCK), args, 0, c->stream));
    }
  };
  if (block >= 1024)
    run(bicg_fused_kernel<T, V, PC, 1024, true>, bicg_fused_kernel<T, V, PC, 1024, false>, 1024);
  else if (block >= 512)
    run(bicg_fused_kernel<T, V, PC, 512, true>, bicg_fused_kernel<T, V, PC, 512, false>, 512);
  else
    run(bicg_fused_kernel<T, V, PC, 256, true>, bicg_fused_kernel<T, V, PC, 256, false>, 256);
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
