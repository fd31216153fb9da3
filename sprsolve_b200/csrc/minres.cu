// minres.cu -- device-resident MINRES (MinRes::solve / precond_solve, src/minres.rs:31-172,
// :178-341) and CSMINRES for complex-symmetric systems (CSMinRes::solve, src/cs_minres.rs:29-158).
//
// Lanczos (Saunders for CS) + Givens recurrences exactly as written in the reference, including
// the pointer rotation of the Lanczos / direction vectors (done on the host: it is data
// independent), the 0-based iteration index and the estimated residual prod |s|.  Fusion:
//   SpMV : v_new = A q  (CS: A conj(q), the conjugate is applied in the gather, tvec is never
//          materialised) with the epilogue alpha = conj(q) . v_new            (minres.rs:116, cs:99-103)
//   KM1  : v_new -= beta v_old ; v_new -= alpha v ; ||v_new||^2               (minres.rs:117-120)
//          [Jacobi: w_new = M v_new and <v_new, w_new> in the same pass,      minres.rs:275-278]
//   S    : beta_new, Givens rotation, residual estimate, convergence          (minres.rs:120-148,164-168)
//   KM2  : v_new *= 1/beta_new ; p = (q - r2 p_old - r3 p_oold) r1_inv ; x += tau p   (:121,151-162)
// n-vector streams per iteration: SpMV epilogue 1R, KM1 3R+1W, KM2 5R+3W = 13.
#include "finalize.cuh"
#include "solver.cuh"

namespace spb {

template <typename T>
struct MinresState {
  StateHead h;
  T c, c_old, eta, alpha, nalpha, nbeta, nr2, nr3, tau;
  real_t<T> s, s_old, beta, beta_new, beta_one, res_norm, threshold, rhs_norm, tol, inv_beta, r1_inv;  // T::Real
};

__device__ __forceinline__ void mr_hist_put(StateHead& h, double* hist, long long cap, long long k, double v) {
  if (hist && k < cap) hist[k] = v;
  if (k + 1 > h.hist_len) h.hist_len = k + 1;
}

template <typename T>
__global__ void mr_s_rhs(MinresState<T>* st, const scal2* red) {
  using R = real_t<T>;
  const R rhs_norm = sqrt_r((R)red[0].re);  // minres.rs:51
  st->rhs_norm = rhs_norm;
  st->threshold = st->tol * rhs_norm;       // :57
  if (rhs_norm <= eps_of<T>()) {            // :52-56
    st->h.status = DS_ZERO_RHS;
    st->h.res_iters = 0;
    st->h.res_resid = (double)rhs_norm;
  }
}

// precond: b2 = <v_new, w_new>; validity test minres.rs:236-244 / :279-287
template <typename T>
__device__ __forceinline__ bool beta_from_precond(const scal2& b2, real_t<T>* beta_new) {
  using R = real_t<T>;
  const R re = (R)b2.re, im = (R)b2.im;
  if (re < eps_of<T>() || im > eps_of<T>() * re) return false;
  *beta_new = sqrt_r(re);
  return true;
}

template <typename T>
__global__ void mr_s_init(MinresState<T>* st, const scal2* red, const scal2* b2, int precond) {
  if (st->h.status != DS_RUNNING) return;
  using R = real_t<T>;
  st->res_norm = sqrt_r((R)red[0].re);  // :81 / :231
  R beta_new;
  if (precond) {
    if (!beta_from_precond<T>(*b2, &beta_new)) {
      st->h.status = DS_INVALID_PRECOND;
      st->h.res_iters = 0;
      return;
    }
  } else {
    beta_new = st->res_norm;  // :82
  }
  st->beta_new = beta_new;
  st->beta_one = beta_new;           // :83 / :246
  st->inv_beta = (R)1 / beta_new;    // :84 / :248
  st->c = one_of<T>();               // :60-64
  st->c_old = one_of<T>();
  st->s = (R)0;
  st->s_old = (R)0;
  st->eta = one_of<T>();
}

template <typename T>
__device__ __forceinline__ void mr_s_alpha_body(MinresState<T>* st, const scal2* red) {
  if (st->h.status != DS_RUNNING) return;
  st->beta = st->beta_new;               // :91
  st->alpha = from_scal2<T>(red[0]);     // :116
  st->nalpha = neg(st->alpha);
  st->nbeta = from_real<T>(-st->beta);   // T::from_real(-beta), :117
}
// fused into the kernel that finishes alpha = <q, A q> (finalize.cuh)
template <typename T>
struct MrAlphaTail {
  MinresState<T>* st;
  const scal2* red;
  __device__ __forceinline__ void operator()() const { mr_s_alpha_body(st, red); }
};

template <typename T, bool CS>
__device__ __forceinline__ void mr_s_givens_body(MinresState<T>* st, const scal2* red, const scal2* b2, int precond,
                                                 long long its, double* hist, long long cap) {
  if (st->h.status != DS_RUNNING) return;
  using R = real_t<T>;
  R beta_new;
  if (precond) {
    if (!beta_from_precond<T>(*b2, &beta_new)) {  // :279-287
      st->h.status = DS_INVALID_PRECOND;
      st->h.res_iters = its;
      return;
    }
  } else {
    beta_new = sqrt_r((R)red[0].re);  // :120
  }
  st->beta_new = beta_new;
  st->inv_beta = (R)1 / beta_new;  // :121 / :289
  const R beta = st->beta, s = st->s, s_old = st->s_old;
  const T c = st->c, c_old = st->c_old, alpha = st->alpha;
  // Givens rotation, minres.rs:132-148 ; cs_minres.rs:119-134 (conjugations differ)
  const R r3 = s_old * beta;
  const T tr = CS ? mul_real(conj_of(c_old), beta) : mul_real(c_old, beta);
  const T r2 = add(mul_real(alpha, s), mul(c, tr));
  const T r1_hat = CS ? sub(mul(conj_of(c), alpha), mul_real(tr, s)) : sub(mul(c, alpha), mul_real(tr, s));
  const R r1_inv = (R)1 / sqrt_r(square(r1_hat) + beta_new * beta_new);
  st->c_old = c;
  st->s_old = s;
  const T c_new = CS ? mul_real(conj_of(r1_hat), r1_inv) : mul_real(r1_hat, r1_inv);
  const R s_new = beta_new * r1_inv;
  st->c = c_new;
  st->s = s_new;
  st->nr2 = neg(r2);                 // axpy(-r2, p_old, p), :158
  st->nr3 = from_real<T>(-r3);       // axpy(T::from_real(-r3), p_oold, p), :159
  st->r1_inv = r1_inv;               // rscale(r1_inv, p), :160
  st->tau = mul_real(mul(c_new, st->eta), st->beta_one);  // :162
  st->res_norm *= fabs_r(s_new);     // :164
  mr_hist_put(st->h, hist, cap, its, (double)(st->res_norm / st->rhs_norm));
  if (st->res_norm < st->threshold) {  // :165-167 (the x update of this iteration still runs)
    st->h.status = DS_OK;
    st->h.res_iters = its;
    st->h.res_resid = (double)(st->res_norm / st->rhs_norm);
    return;
  }
  st->eta = mul_real(st->eta, -s_new);  // :168
}
template <typename T, bool CS>
__global__ void mr_s_givens(MinresState<T>* st, const scal2* red, const scal2* b2, int precond, long long its, double* hist,
                            long long cap) {
  mr_s_givens_body<T, CS>(st, red, b2, precond, its, hist, cap);
}
// fused into the kernel that finishes ||v_new||^2 (and <v_new, w_new> with Jacobi)
template <typename T, bool CS>
struct MrGivensTail {
  MinresState<T>* st;
  const scal2* red;
  const scal2* b2;
  int precond;
  long long its;
  double* hist;
  long long cap;
  __device__ __forceinline__ void operator()() const { mr_s_givens_body<T, CS>(st, red, b2, precond, its, hist, cap); }
};

// v_new = rhs - A x (A x is in v_old) ; ||v_new||^2 ; v = p_old = p = 0      (minres.rs:77-88)
template <typename T, typename V, bool JACOBI>
__global__ void __launch_bounds__(kVecThreads)
mr_k_init(const MinresState<T>* st, int64_t n, const T* rhs, const T* v_old, T* v_new, T* v, T* p_old, T* p,
          T* w_new, const V* dinv, Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
  if (st->h.status == DS_RUNNING) {
    const T m1 = neg(one_of<T>()), z = zero_of<T>();
    SPB_GRID_STRIDE(i, n) {
      const T vn = add(rhs[i], mul(v_old[i], m1));  // axpy(-1, v_old, v_new), :80
      v_new[i] = vn;
      acc_sq(e0, vn);                               // :81
      if (JACOBI) {
        const T wn = mul_diag(vn, dinv[i]);         // w_new = M v_new, :233
        w_new[i] = wn;
        acc_prod(e1, conj_of(vn), wn);              // :235
      }
      v[i] = z;
      p_old[i] = z;
      p[i] = z;
    }
  }
  write_partials(e0, e1, partials);
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads)
mr_k_scale(const MinresState<T>* st, int64_t n, T* v_new, T* w_new) {
  if (st->h.status != DS_RUNNING) return;
  const real_t<T> ts = st->inv_beta;
  SPB_GRID_STRIDE(i, n) {
    v_new[i] = mul_real(v_new[i], ts);              // rscale, :84 / :249
    if (w_new) w_new[i] = mul_real(w_new[i], ts);   // :250
  }
}

template <typename T, typename V, bool JACOBI>
__global__ void __launch_bounds__(kVecThreads)
mr_k1(const MinresState<T>* st, int64_t n, T* v_new, const T* v_old, const T* v, T* w_new, const V* dinv,
      Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>(), e1 = zero_of<Acc<T>>();
  if (st->h.status == DS_RUNNING) {
    const T nbeta = st->nbeta, nalpha = st->nalpha;
    SPB_GRID_STRIDE(i, n) {
      T vn = add(v_new[i], mul(v_old[i], nbeta));  // :117
      vn = add(vn, mul(v[i], nalpha));             // :118
      v_new[i] = vn;
      acc_sq(e0, vn);                              // :120
      if (JACOBI) {
        const T wn = mul_diag(vn, dinv[i]);        // :276
        w_new[i] = wn;
        acc_prod(e1, conj_of(vn), wn);             // :278
      }
    }
  }
  write_partials(e0, e1, partials);
}

template <typename T, bool CS>
__global__ void __launch_bounds__(kVecThreads)
mr_k2(const MinresState<T>* st, long long its, int64_t n, const T* q, const T* p_old, const T* p_oold, T* p,
      T* x, T* v_new, T* w_new) {
  const int status = st->h.status;
  if (!(status == DS_RUNNING || (status == DS_OK && st->h.res_iters == its))) return;
  const T nr2 = st->nr2, nr3 = st->nr3, tau = st->tau;
  const real_t<T> r1_inv = st->r1_inv, ts = st->inv_beta;
  SPB_GRID_STRIDE(i, n) {
    v_new[i] = mul_real(v_new[i], ts);             // rscale(1/beta_new, v_new), :121 / :290
    if (w_new) w_new[i] = mul_real(w_new[i], ts);  // :291
    T pi = CS ? conj_of(q[i]) : q[i];              // p = q (:156) ; CS: p = conj(q_k) (cs:99,142)
    pi = add(pi, mul(p_old[i], nr2));              // :158
    pi = add(pi, mul(p_oold[i], nr3));             // :159
    pi = mul_real(pi, r1_inv);                     // :160
    p[i] = pi;
    x[i] = add(x[i], mul(pi, tau));                // :162
  }
}

template <typename T>
struct MinRes : spb_solver {
  bool cs;
  DevBuf ws;  // 8 n T (src/minres.rs:24) / 7 n for CS (src/cs_minres.rs:22)
  DevBuf partials, red, red2, state, hist_d;
  MinRes(spb_op* A_, int64_t size_, bool cs_) : cs(cs_) {
    A = A_;
    ctx = A_->ctx;
    kind = cs_ ? 2 : 1;
    dtype = ScalarTraits<T>::dtype;
    size = size_;
    ws.alloc(sizeof(T) * (cs ? 7 : 8) * (size_t)std::max<int64_t>(size, 1));
    SPB_CUDA(cudaMemsetAsync(ws.p, 0, ws.bytes, ctx->stream));  // vec![T::zero(); size*8]
    partials.alloc(sizeof(Acc<T>) * 2 * (size_t)(vec_max_grid(ctx) + 1));
    red.alloc(sizeof(scal2) * 2);
    red2.alloc(sizeof(scal2) * 2);
    state.alloc(sizeof(MinresState<T>));
  }
  int solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double tol, int64_t* iters,
                double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) override;
};

template <typename T>
int MinRes<T>::solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double tol,
                         int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                         int64_t* hist_len) {
  Ctx* c = ctx;
  const int64_t n = size;
  const T* rhs = (const T*)d_rhs;
  T* x = (T*)d_x;
  if (A->kind != OP_CSR) SPB_FAIL(SPB_INVALID_ARG, "MINRES operator must be a CSR matrix");
  auto* Am = static_cast<CsrMat<T>*>(A);
  if (Am->n_local != n) {
    set_last_error("Input vec dimension doesn't match the matrix size");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  if (cs && M) SPB_FAIL(SPB_INVALID_ARG, "CSMinRes has no preconditioned variant (src/cs_minres.rs)");
  if (M && M->n_local != n) SPB_FAIL(SPB_DIM_MISMATCH, "preconditioner dimension mismatch");
  const PcMode pcm = pc_mode_of<T>(M);
  const bool precond = pcm != PCM_NONE;
  const bool jac = pcm == PCM_JACOBI || pcm == PCM_JACOBI_REAL;
  const void* dinv = jac ? static_cast<DiagOp<T>*>(M)->dinv.p : nullptr;
  T* w0 = bufptr<T>(ws);
  T *v_old = w0, *v_new = w0 + n, *v = w0 + 2 * n, *p_old = w0 + 3 * n, *p_oold = w0 + 4 * n, *p = w0 + 5 * n;
  T *w = precond ? w0 + 6 * n : nullptr, *w_new = precond ? w0 + 7 * n : nullptr;
  auto* st = bufptr<MinresState<T>>(state);
  scal2* redp = bufptr<scal2>(red);
  scal2* red2p = bufptr<scal2>(red2);
  Acc<T>* parts = bufptr<Acc<T>>(partials);
  const int grid = vec_grid(c, n);
  // the kernels that carry partial sums (40-48 registers): one wave of resident CTAs (vecops.cuh: vec_grid_resident)
  int grid_r = vec_grid_resident(c, n, mr_k1<T, T, true>);
  grid_r = std::min(grid_r, vec_grid_resident(c, n, mr_k1<T, T, false>));
  grid_r = std::min(grid_r, vec_grid_resident(c, n, mr_k_init<T, T, true>));
  grid_r = std::min(grid_r, vec_grid_resident(c, n, mr_k_init<T, T, false>));
  const long long cap = hist ? std::min<int64_t>(hist_cap, max_iter) : 0;
  double* hd = nullptr;
  if (cap > 0) {
    hist_d.ensure(sizeof(double) * cap);
    hd = bufptr<double>(hist_d);
  }
  MinresState<T> init;
  memset(&init, 0, sizeof(init));
  init.h.status = DS_RUNNING;
  init.tol = (real_t<T>)tol;
  SPB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));

  auto scalar = [&](auto kernel, auto... args) {
    LaunchScope ls(c, FAM_SCALAR);
    kernel<<<1, 1, 0, c->stream>>>(args...);
    check_launch("minres scalar kernel");
  };
  auto reduce_vec = [&]() {
    finalize_allreduce<T>(c, parts, grid_r, redp);
  };
  // b2 = <v_new, w_new> for a generic preconditioner (separate apply + conj_dot)
  auto generic_b2 = [&](T* vn, T* wn) {
    op_apply<T>(M, vn, wn);
    vec_reduce<T>(c, 1, n, vn, wn, parts, red2p, true);  // gated apply; the dot itself is harmless
  };
  const scal2* b2src = jac ? redp + 1 : red2p;

  int rc = SPB_OK;
  Poller poller(c);
  StateHead hh;
  c->gate = nullptr;
  try {
    vec_reduce<T>(c, 2, n, rhs, rhs, parts, redp, true);
    scalar(mr_s_rhs<T>, st, redp);
    poller.post(st);
    poller.drain(&hh);
    if (hh.status == DS_ZERO_RHS) {
      SPB_CUDA(cudaMemsetAsync(x, 0, sizeof(T) * n, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      *iters = 0;
      *resid = hh.res_resid;
      if (hist_len) *hist_len = 0;
      return SPB_OK;
    }
    c->gate = &st->h.status;
    Am->mul(x, v_old, EPI_NONE, nullptr, false);  // v_old = A x, :78
    {
      LaunchScope ls(c, FAM_VEC);
      if (pcm == PCM_JACOBI)
        mr_k_init<T, T, true><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, rhs, v_old, v_new, v, p_old, p, w_new, (const T*)dinv, parts);
      else if (pcm == PCM_JACOBI_REAL)
        mr_k_init<T, real_t<T>, true><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, rhs, v_old, v_new, v, p_old, p, w_new, (const real_t<T>*)dinv, parts);
      else
        mr_k_init<T, T, false><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, rhs, v_old, v_new, v, p_old, p, w_new, (const T*)nullptr, parts);
      check_launch("mr_k_init");
    }
    reduce_vec();
    if (pcm == PCM_GENERIC) generic_b2(v_new, w_new);
    scalar(mr_s_init<T>, st, redp, b2src, precond ? 1 : 0);
    {
      LaunchScope ls(c, FAM_VEC);
      mr_k_scale<T><<<grid, kVecThreads, 0, c->stream>>>(st, n, v_new, w_new);
      check_launch("mr_k_scale");
    }

    bool done = false;
    int64_t its = 0;
    auto handle = [&](const StateHead& h) {
      if (h.status != DS_RUNNING) done = true;
    };
    while (!done) {
      StateHead h;
      if (its < max_iter) {
        const int64_t chunk = std::min<int64_t>(poll, max_iter - its);
        for (int64_t k = 0; k < chunk; ++k, ++its) {
          // pointer rotation, :92-96 / :258-265 / :151-154
          T* vt = v_old;
          v_old = v;
          v = v_new;
          v_new = vt;
          if (precond) std::swap(w, w_new);
          const T* q = precond ? w : v;
          // v_new = A q with alpha = conj(q) . v_new   (CS: v_new = A conj(v), alpha = conj(v) . v_new)
          Am->mul(q, v_new, EPI_DOT_WY, q, cs);
          finalize_reduce_tail<T>(c, bufptr<Acc<T>>(Am->partials), Am->last_partial_blocks, bufptr<scal2>(Am->red), true,
                                  MrAlphaTail<T>{st, bufptr<scal2>(Am->red)});  // + alpha, -beta
          {
            LaunchScope ls(c, FAM_VEC);
            if (pcm == PCM_JACOBI)
              mr_k1<T, T, true><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, v_new, v_old, v, w_new, (const T*)dinv, parts);
            else if (pcm == PCM_JACOBI_REAL)
              mr_k1<T, real_t<T>, true><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, v_new, v_old, v, w_new, (const real_t<T>*)dinv, parts);
            else
              mr_k1<T, T, false><<<grid_r, kVecThreads, 0, c->stream>>>(st, n, v_new, v_old, v, w_new, (const T*)nullptr, parts);
            check_launch("mr_k1");
          }
          if (pcm == PCM_GENERIC) {  // beta^2 = <v_new, M v_new> needs the operator in between: not fused
            reduce_vec();
            generic_b2(v_new, w_new);
            scalar(mr_s_givens<T, false>, st, redp, b2src, 1, (long long)its, hd, cap);
          } else if (cs) {
            finalize_reduce_tail<T>(c, parts, grid_r, redp, true, MrGivensTail<T, true>{st, redp, b2src, 0, (long long)its, hd, cap});
          } else {
            finalize_reduce_tail<T>(c, parts, grid_r, redp, true,
                                    MrGivensTail<T, false>{st, redp, b2src, precond ? 1 : 0, (long long)its, hd, cap});
          }
          T* pt = p_oold;
          p_oold = p_old;
          p_old = p;
          p = pt;
          {
            LaunchScope ls(c, FAM_VEC);
            if (cs)
              mr_k2<T, true><<<grid, kVecThreads, 0, c->stream>>>(st, (long long)its, n, q, p_old, p_oold, p, x, v_new, w_new);
            else
              mr_k2<T, false><<<grid, kVecThreads, 0, c->stream>>>(st, (long long)its, n, q, p_old, p_oold, p, x, v_new, w_new);
            check_launch("mr_k2");
          }
        }
        poller.post(st);
        if (poller.wait_oldest(&h)) handle(h);
      } else {
        if (!poller.drain(&h)) {
          poller.post(st);
          poller.drain(&h);
        }
        handle(h);
        break;
      }
    }
    c->gate = nullptr;
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    poller.post(st);
    poller.drain(&hh);
    if (hh.status == DS_OK) {
      *iters = hh.res_iters;
      *resid = hh.res_resid;
      rc = SPB_OK;
    } else if (hh.status == DS_INVALID_PRECOND) {
      *iters = hh.res_iters;
      rc = SPB_INVALID_PRECOND;
    } else {
      *iters = max_iter;
      rc = SPB_INSUFFICIENT_ITER;
    }
    const int64_t hl = hh.hist_len;
    if (hist_len) *hist_len = hl;
    if (hd && hl > 0)
      SPB_CUDA(cudaMemcpy(hist, hd, sizeof(double) * std::min<int64_t>(hl, cap), cudaMemcpyDeviceToHost));
  } catch (...) {
    c->gate = nullptr;
    throw;
  }
  return rc;
}

spb_solver* make_minres(spb_op* A, int64_t size, bool cs) {
  switch (A->dtype) {
    case SPB_F64: return new MinRes<double>(A, size, cs);
    case SPB_C128: return new MinRes<cplx>(A, size, cs);
    case SPB_F32: return new MinRes<float>(A, size, cs);
    default: return new MinRes<cplxf>(A, size, cs);
  }
}

}  // namespace spb
