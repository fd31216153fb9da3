// gs_solver.cu -- device-resident stationary Gauss-Seidel solver: GaussSeidel::new / solve
// (src/gauss_seidel.rs:15-31, :33-140).  Natural row order and "latest value" semantics are kept
// by the level-scheduled sweep of ops.cu (rows < i read the new iterate, rows > i the old one);
// the per-sweep residual ||A x - b|| is one SpMV plus one fused axpy+norm kernel (:127-133).
// Returns the ABSOLUTE residual and Ok((1, res)) when the first sweep converges (:106-108).
#include "solver.cuh"

namespace spb {

template <typename R>
struct GsState {
  StateHead h;
  R tol2, eps;  // T::Real
};

template <typename R>
__global__ void gs_s_bnorm(GsState<R>* st, const scal2* red) {
  st->tol2 = st->eps * sqrt_r((R)red[0].re);  // src/gauss_seidel.rs:83,87
}

template <typename R>
__global__ void gs_s_check(GsState<R>* st, const scal2* red, long long it, double* hist, long long cap) {
  if (st->h.status != DS_RUNNING) return;
  const R res = sqrt_r((R)red[0].re);  // :104 / :133
  if (hist && it < cap) hist[it] = (double)res;
  if (it + 1 > st->h.hist_len) st->h.hist_len = it + 1;
  st->h.its = it;
  if (res <= st->tol2) {  // :106-108 / :135-137
    st->h.status = DS_OK;
    st->h.res_iters = it == 0 ? 1 : it;
    st->h.res_resid = (double)res;
  }
}

template <typename T>
__global__ void __launch_bounds__(kVecThreads)
gs_k_resid(const GsState<real_t<T>>* st, int64_t n, const T* rhs, T* res, Acc<T>* partials) {
  Acc<T> e0 = zero_of<Acc<T>>();
  if (st->h.status == DS_RUNNING) {
    const T m1 = neg(one_of<T>());
    SPB_GRID_STRIDE(i, n) {
      const T ri = add(res[i], mul(rhs[i], m1));  // axpy(-1, rhs, r), :97 / :131
      res[i] = ri;
      acc_sq(e0, ri);                             // norm2, :104 / :133
    }
  }
  write_partials(e0, zero_of<Acc<T>>(), partials);
}

template <typename T>
struct GaussSeidelSolver : spb_solver {
  GsOp<T>* gs = nullptr;
  DevBuf res, xalt, partials, red, state, hist_d;
  explicit GaussSeidelSolver(spb_op* A_, double omega) {
    A = A_;
    ctx = A_->ctx;
    kind = 3;
    dtype = ScalarTraits<T>::dtype;
    size = A_->n_local;
    gs = gs_create<T>(static_cast<CsrMat<T>*>(A_), SPB_GS_FORWARD, omega);
    const size_t n1 = (size_t)std::max<int64_t>(size, 1);
    res.alloc(sizeof(T) * n1);   // workspace[0..n]   (src/gauss_seidel.rs:29)
    xalt.alloc(sizeof(T) * n1);
    partials.alloc(sizeof(Acc<T>) * 2 * (size_t)(vec_max_grid(ctx) + 1));
    red.alloc(sizeof(scal2) * 2);
    state.alloc(sizeof(GsState<real_t<T>>));
  }
  ~GaussSeidelSolver() override { delete gs; }
  int solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double eps, int64_t* iters,
                double* resid, double* hist, int64_t hist_cap, int64_t* hist_len) override;
};

template <typename T>
int GaussSeidelSolver<T>::solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double eps,
                                    int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                                    int64_t* hist_len) {
  Ctx* c = ctx;
  const int64_t n = size;
  if (M) SPB_FAIL(SPB_INVALID_ARG, "GaussSeidel::solve takes no preconditioner");
  if (hist_len) *hist_len = 0;
  if (max_iter == 0) {  // src/gauss_seidel.rs:52-54
    *iters = 0;
    return SPB_INSUFFICIENT_ITER;
  }
  if (gs->bad_row >= 0) {  // :72-78
    *iters = gs->bad_row;
    return SPB_ZERO_DIAGONAL;
  }
  auto* Am = static_cast<CsrMat<T>*>(A);
  const T* rhs = (const T*)d_rhs;
  T* xbuf[2] = {(T*)d_x, bufptr<T>(xalt)};
  using R = real_t<T>;
  auto* st = bufptr<GsState<R>>(state);
  scal2* redp = bufptr<scal2>(red);
  Acc<T>* parts = bufptr<Acc<T>>(partials);
  T* resv = bufptr<T>(res);
  const int grid = vec_grid(c, n);
  const long long cap = hist ? std::min<int64_t>(hist_cap, max_iter) : 0;
  double* hd = nullptr;
  if (cap > 0) {
    hist_d.ensure(sizeof(double) * cap);
    hd = bufptr<double>(hist_d);
  }
  GsState<R> init;
  memset(&init, 0, sizeof(init));
  init.h.status = DS_RUNNING;
  init.eps = (R)eps;
  SPB_CUDA(cudaMemcpyAsync(st, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  int rc = SPB_OK;
  Poller poller(c);
  c->gate = nullptr;
  try {
    vec_reduce<T>(c, 2, n, rhs, rhs, parts, redp);  // ||b||^2, :83
    {
      LaunchScope ls(c, FAM_SCALAR);
      gs_s_bnorm<R><<<1, 1, 0, c->stream>>>(st, redp);
      check_launch("gs_s_bnorm");
    }
    c->gate = &st->h.status;
    bool done = false;
    int64_t it = 0;
    StateHead h;
    while (!done) {
      if (it < max_iter) {
        const int64_t chunk = std::min<int64_t>(poll, max_iter - it);
        for (int64_t k = 0; k < chunk; ++k, ++it) {
          const T* xo = xbuf[it & 1];
          T* xn = xbuf[(it + 1) & 1];
          gs_solver_sweep<T>(gs, rhs, xo, xn);               // :60-86 / :111-125
          Am->mul(xn, resv, EPI_NONE, nullptr, false);       // :90 / :128
          {
            LaunchScope ls(c, FAM_VEC);
            gs_k_resid<T><<<grid, kVecThreads, 0, c->stream>>>(st, n, rhs, resv, parts);
            check_launch("gs_k_resid");
          }
          finalize_partials<T>(c, parts, grid, redp);
          {
            LaunchScope ls(c, FAM_SCALAR);
            gs_s_check<R><<<1, 1, 0, c->stream>>>(st, redp, (long long)it, hd, cap);
            check_launch("gs_s_check");
          }
        }
        poller.post(st);
        if (poller.wait_oldest(&h) && h.status != DS_RUNNING) done = true;
      } else {
        break;
      }
    }
    c->gate = nullptr;
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    poller.drain(&h);
    poller.post(st);
    poller.drain(&h);
    // sweep h.its was the last one executed: its result lives in xbuf[(h.its + 1) & 1]
    const int64_t last = h.its;
    if (((last + 1) & 1) == 1)
      SPB_CUDA(cudaMemcpyAsync(xbuf[0], xbuf[1], sizeof(T) * n, cudaMemcpyDeviceToDevice, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    if (h.status == DS_OK) {
      *iters = h.res_iters;
      *resid = h.res_resid;
      rc = SPB_OK;
    } else {
      *iters = max_iter;
      rc = SPB_INSUFFICIENT_ITER;
    }
    if (hist_len) *hist_len = h.hist_len;
    if (hd && h.hist_len > 0)
      SPB_CUDA(cudaMemcpy(hist, hd, sizeof(double) * std::min<int64_t>(h.hist_len, cap), cudaMemcpyDeviceToHost));
  } catch (...) {
    c->gate = nullptr;
    throw;
  }
  return rc;
}

spb_solver* make_gauss_seidel(spb_op* A, double omega) {
  if (A->kind != OP_CSR) {
    set_last_error("Not in CSR format");  // src/gauss_seidel.rs:22-26
    throw SpbError{SPB_INCOMPATIBLE_FORMAT};
  }
  switch (A->dtype) {
    case SPB_F64: return new GaussSeidelSolver<double>(A, omega);
    case SPB_C128: return new GaussSeidelSolver<cplx>(A, omega);
    case SPB_F32: return new GaussSeidelSolver<float>(A, omega);
    default: return new GaussSeidelSolver<cplxf>(A, omega);
  }
}

}  // namespace spb
