// ops.cu -- Jacobi (DiagPrecond, src/precond.rs:6-63) and the level-scheduled Gauss-Seidel sweep
// (sweep body src/gauss_seidel.rs:111-125; first-sweep diagonal cache / zero-diagonal test
// :60-86).  The GS sweep keeps the reference's sequential update order: the dependency DAG of the
// strictly-lower (forward) / strictly-upper (backward) pattern is cut into levels at analysis
// time; inside a level every row is independent; one thread per row accumulates sigma in CSR
// order, so each x_i is bit-identical to the sequential loop.  One persistent cooperative kernel
// walks all levels with a grid-wide barrier in between (no kernel launch per level).
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ops.cuh"
#include "vecops.cuh"

namespace spb {

// ---------------------------------------------------------------- Jacobi
template <typename V>
__global__ void reciprocal_kernel(int64_t n, V* d);
template <>
__global__ void reciprocal_kernel<double>(int64_t n, double* d) {
  SPB_GRID_STRIDE(i, n) d[i] = 1.0 / d[i];  // V::one() / *v, src/precond.rs:23
}
template <>
__global__ void reciprocal_kernel<cplx>(int64_t n, cplx* d) {
  SPB_GRID_STRIDE(i, n) d[i] = divi(cplx{1.0, 0.0}, d[i]);
}
template <>
__global__ void reciprocal_kernel<float>(int64_t n, float* d) {
  SPB_GRID_STRIDE(i, n) d[i] = 1.0f / d[i];
}
template <>
__global__ void reciprocal_kernel<cplxf>(int64_t n, cplxf* d) {
  SPB_GRID_STRIDE(i, n) d[i] = divi(cplxf{1.0f, 0.0f}, d[i]);
}

template <typename T, typename V>
__global__ void __launch_bounds__(kVecThreads) diag_apply_kernel(int64_t n, const V* dinv, const T* in, T* out, const int* gate, int gate_value) {
  if (gate && *gate != gate_value) return;
  SPB_GRID_STRIDE(i, n) out[i] = mul_diag(in[i], dinv[i]);  // src/precond.rs:49-51
}

template <typename T>
DiagOp<T>* diag_from_host(Ctx* ctx, int diag_dtype, const void* diag, int64_t n) {
  auto* op = new DiagOp<T>();
  try {
    op->ctx = ctx;
    op->kind = OP_DIAG;
    op->dtype = ScalarTraits<T>::dtype;
    op->n_global = op->n_local = n;
    op->real_diag = ScalarTraits<T>::is_complex && diag_dtype == ScalarTraits<real_t<T>>::dtype;
    const size_t esz = op->real_diag ? sizeof(real_t<T>) : sizeof(T);
    op->dinv.alloc(esz * (size_t)std::max<int64_t>(n, 1));
    if (n) SPB_CUDA(cudaMemcpyAsync(op->dinv.p, diag, esz * n, cudaMemcpyHostToDevice, ctx->stream));
    if (n) {
      LaunchScope ls(ctx, FAM_PRECOND);
      if (op->real_diag || !ScalarTraits<T>::is_complex)
        reciprocal_kernel<real_t<T>><<<vec_grid(ctx, n), kVecThreads, 0, ctx->stream>>>(n, bufptr<real_t<T>>(op->dinv));
      else
        reciprocal_kernel<T><<<vec_grid(ctx, n), kVecThreads, 0, ctx->stream>>>(n, bufptr<T>(op->dinv));
      check_launch("reciprocal_kernel");
    }
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    delete op;
    throw;
  }
  return op;
}

template <typename T, typename IP>
__global__ void csr_diag_kernel(const IP* indptr, const int* cols, const T* vals, int64_t n, T* diag,
                                unsigned long long* bad_row) {
  SPB_GRID_STRIDE(r, n) {
    T d = zero_of<T>();
    bool found = false;
    for (IP k = indptr[r]; k < indptr[r + 1]; ++k)
      if (cols[k] == (int)r) {
        d = vals[k];
        found = true;
      }
    diag[r] = d;
    // src/gauss_seidel.rs:72-78: missing diagonal or |d|^2 < eps
    if (bad_row && (!found || square(d) < eps_of<T>())) atomicMin(bad_row, (unsigned long long)r);
  }
}

template <typename T>
static void csr_diag_impl(CsrMat<T>* A, T* d_diag, unsigned long long* d_bad) {
  Ctx* c = A->ctx;
  const int64_t n = A->n_local;
  if (n == 0) return;
  LaunchScope ls(c, FAM_PRECOND);
  if (A->ip64)
    csr_diag_kernel<T, int64_t><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(bufptr<int64_t>(A->indptr), bufptr<int>(A->cols), bufptr<T>(A->vals), n, d_diag, d_bad);
  else
    csr_diag_kernel<T, int32_t><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(bufptr<int32_t>(A->indptr), bufptr<int>(A->cols), bufptr<T>(A->vals), n, d_diag, d_bad);
  check_launch("csr_diag_kernel");
}

template <typename T>
void csr_diagonal(CsrMat<T>* A, T* d_diag) {
  csr_diag_impl(A, d_diag, nullptr);
}

template <typename T>
DiagOp<T>* diag_from_csr(CsrMat<T>* A) {
  Ctx* ctx = A->ctx;
  auto* op = new DiagOp<T>();
  try {
    op->ctx = ctx;
    op->kind = OP_DIAG;
    op->dtype = ScalarTraits<T>::dtype;
    op->n_global = A->n_global;
    op->n_local = A->n_local;
    op->row_begin = A->row_begin;
    const int64_t n = A->n_local;
    op->dinv.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
    csr_diag_impl(A, bufptr<T>(op->dinv), nullptr);
    if (n) {
      LaunchScope ls(ctx, FAM_PRECOND);
      reciprocal_kernel<T><<<vec_grid(ctx, n), kVecThreads, 0, ctx->stream>>>(n, bufptr<T>(op->dinv));
      check_launch("reciprocal_kernel");
    }
    SPB_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    delete op;
    throw;
  }
  return op;
}

template <typename T>
void diag_apply(DiagOp<T>* M, const T* in, T* out) {
  Ctx* c = M->ctx;
  const int64_t n = M->n_local;
  if (n == 0) return;
  LaunchScope ls(c, FAM_PRECOND);
  if (M->real_diag)
    diag_apply_kernel<T, real_t<T>><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(n, bufptr<real_t<T>>(M->dinv), in, out, c->gate, c->gate_value);
  else
    diag_apply_kernel<T, T><<<vec_grid(c, n), kVecThreads, 0, c->stream>>>(n, bufptr<T>(M->dinv), in, out, c->gate, c->gate_value);
  check_launch("diag_apply_kernel");
}

// ---------------------------------------------------------------- Gauss-Seidel
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cplx ld_cg(const cplx* p) {
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return cplx{v.x, v.y};
}
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ cplxf ld_cg(const cplxf* p) {
  const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  return cplxf{v.x, v.y};
}

// All CTAs are co-resident (cooperative launch), so a counter barrier cannot deadlock.
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1ULL);
    while (*((volatile unsigned long long*)counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T, typename IP>
struct GsArgs {
  const IP* indptr;
  const int* cols;
  const T* vals;
  const T* diag;
  const int* level_ptr;
  const int* rows;
  int nlevels;
  const T* rhs;
  const T* lo_src;  // values for columns < row (null: skip)
  const T* hi_src;  // values for columns > row (null: skip -- they multiply exact zeros)
  T* out;
  unsigned long long* counter;
  const int* gate;
  int gate_value;
  // relaxed sweep (SOR / SSOR(omega); not in the reference): x_i <- (1 - omega) xold_i + omega g_i, computed as
  // mul_real(xold_i, 1 - omega) + mul_real(g_i, omega); xold null = a sweep from zero.  relaxed == 0: plain sweep.
  int relaxed;
  real_t<T> omega;
  const T* xold;
};

template <typename T, typename IP>
__global__ void __launch_bounds__(256) gs_sweep_kernel(const GsArgs<T, IP> a) {
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  unsigned long long target = 0;
  if (a.gate && *a.gate != a.gate_value) return;
  for (int lvl = 0; lvl < a.nlevels; ++lvl) {
    const int b = a.level_ptr[lvl], e = a.level_ptr[lvl + 1];
    for (int64_t i = b + gtid; i < e; i += gthreads) {
      const int row = a.rows[i];
      T sigma = zero_of<T>();
      for (IP k = a.indptr[row]; k < a.indptr[row + 1]; ++k) {
        const int col = a.cols[k];
        // src/gauss_seidel.rs:113-118: sigma += val * x[col] for col != row, CSR order
        if (col < row) {
          if (a.lo_src) sigma = add(sigma, mul(a.vals[k], ld_cg(a.lo_src + col)));
        } else if (col > row) {
          if (a.hi_src) sigma = add(sigma, mul(a.vals[k], ld_cg(a.hi_src + col)));
        }
      }
      T g = divi(sub(a.rhs[row], sigma), a.diag[row]);  // :123
      if (a.relaxed) {
        const T xo = a.xold ? ld_cg(a.xold + row) : zero_of<T>();
        g = add(mul_real(xo, (real_t<T>)1 - a.omega), mul_real(g, a.omega));
      }
      a.out[row] = g;
    }
    if (lvl + 1 < a.nlevels) grid_barrier(a.counter, target);
  }
}

static void build_levels(Ctx* c, int64_t n, const std::vector<int64_t>& ip, const std::vector<int>& cols,
                         bool lower, LevelSched& out) {
  std::vector<int> level(n, 0);
  int maxl = -1;
  auto visit = [&](int64_t i) {
    int l = 0;
    for (int64_t k = ip[i]; k < ip[i + 1]; ++k) {
      const int j = cols[k];
      if (lower ? (j < i) : (j > i)) l = std::max(l, level[j] + 1);
    }
    level[i] = l;
    maxl = std::max(maxl, l);
  };
  if (lower)
    for (int64_t i = 0; i < n; ++i) visit(i);
  else
    for (int64_t i = n - 1; i >= 0; --i) visit(i);
  const int nl = maxl + 1;
  std::vector<int> ptr(nl + 1, 0), rows(n);
  for (int64_t i = 0; i < n; ++i) ptr[level[i] + 1]++;
  for (int l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
  std::vector<int> cur(ptr.begin(), ptr.end() - 1);
  for (int64_t i = 0; i < n; ++i) rows[cur[level[i]]++] = (int)i;
  out.nlevels = nl;
  out.level_ptr_host = ptr;
  out.level_ptr.alloc(sizeof(int) * (nl + 1));
  out.rows.alloc(sizeof(int) * (size_t)std::max<int64_t>(n, 1));
  SPB_CUDA(cudaMemcpyAsync(out.level_ptr.p, ptr.data(), sizeof(int) * (nl + 1), cudaMemcpyHostToDevice, c->stream));
  if (n) SPB_CUDA(cudaMemcpyAsync(out.rows.p, rows.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
}

template <typename T>
GsOp<T>* gs_create(CsrMat<T>* A, int mode, double omega) {
  Ctx* c = A->ctx;
  if (!(omega > 0.0 && omega < 2.0)) SPB_FAIL(SPB_INVALID_ARG, "relaxation factor must lie in (0, 2)");
  if (c->world() > 1)
    SPB_FAIL(SPB_INVALID_ARG, "Gauss-Seidel in natural order does not partition across GPUs (replicas only)");
  if (mode != SPB_GS_FORWARD && mode != SPB_GS_SYMMETRIC) SPB_FAIL(SPB_INVALID_ARG, "bad gs mode");
  auto* op = new GsOp<T>();
  try {
    op->ctx = c;
    op->kind = OP_GS;
    op->dtype = ScalarTraits<T>::dtype;
    op->n_global = op->n_local = A->n_local;
    op->A = A;
    op->mode = mode;
    op->omega = omega;
    const bool relaxed = (real_t<T>)omega != (real_t<T>)1;
    const int64_t n = A->n_local;
    op->diag.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
    op->tmp.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
    op->barrier.alloc(sizeof(unsigned long long) * 2);
    unsigned long long init = ~0ULL;
    SPB_CUDA(cudaMemcpyAsync(bufptr<unsigned long long>(op->barrier) + 1, &init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    csr_diag_impl(A, bufptr<T>(op->diag), bufptr<unsigned long long>(op->barrier) + 1);
    unsigned long long bad = ~0ULL;
    SPB_CUDA(cudaMemcpyAsync(&bad, bufptr<unsigned long long>(op->barrier) + 1, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
    // host copy of the pattern for the level analysis
    std::vector<int64_t> ip(n + 1);
    std::vector<int> cols((size_t)A->nnz);
    if (A->ip64) {
      SPB_CUDA(cudaMemcpyAsync(ip.data(), A->indptr.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
      std::vector<int> ip32(n + 1);
      SPB_CUDA(cudaMemcpyAsync(ip32.data(), A->indptr.p, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      for (int64_t i = 0; i <= n; ++i) ip[i] = ip32[i];
    }
    if (A->nnz) SPB_CUDA(cudaMemcpyAsync(cols.data(), A->cols.p, sizeof(int) * A->nnz, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    op->bad_row = bad == ~0ULL ? -1 : (int64_t)bad;
    {  // block-wavefront schedules (gs_wave.cu) -- the fast path.  The two directions are independent
       // host analyses of the same pattern: a second thread builds the backward one.
      std::vector<T> vals((size_t)A->nnz);
      if (A->nnz) SPB_CUDA(cudaMemcpyAsync(vals.data(), A->vals.p, sizeof(T) * A->nnz, cudaMemcpyDeviceToHost, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      int bwd_status = SPB_OK;
      std::thread bwd_thread;
      // (the relaxed sweeps run on the level-scheduled kernel: no wavefront schedules)
      if (mode == SPB_GS_SYMMETRIC && !relaxed)
        bwd_thread = std::thread([&]() {
          try {
            SPB_CUDA(cudaSetDevice(c->device));
            wave_build<T>(A, ip, cols, vals, true, op->wbwd);
          } catch (const SpbError& e) {
            bwd_status = e.status;
          } catch (...) {
            bwd_status = SPB_CUDA_ERROR;
          }
        });
      int fwd_status = SPB_OK;
      try {
        if (!relaxed) wave_build<T>(A, ip, cols, vals, false, op->wfwd);
      } catch (const SpbError& e) {
        fwd_status = e.status;
      } catch (...) {  // e.g. std::bad_alloc from the host vectors: never unwind past a joinable thread
        fwd_status = SPB_CUDA_ERROR;
        set_last_error("Gauss-Seidel analysis failed (host allocation?)");
      }
      if (bwd_thread.joinable()) bwd_thread.join();
      if (fwd_status != SPB_OK) throw SpbError{fwd_status};
      if (bwd_status != SPB_OK) throw SpbError{bwd_status};
      if (getenv("SPB_GS_STATS") && op->wfwd.ok) {
        op->wave_stats.alloc(sizeof(long long) * 4 * (size_t)op->wfwd.nblocks);
        SPB_CUDA(cudaMemset(op->wave_stats.p, 0, op->wave_stats.bytes));
      }
      // launch shape of the sweeps (thread-block clusters of 16 / 8 CTAs or plain CTAs): timed once, here, on scratch
      // vectors -- inside a solver the launch queue is gated and nothing could be timed (csrc/gs_wave.cu: wave_sweep)
      if ((op->wfwd.ok || op->wbwd.ok) && n > 0) {
        DevBuf tr, to, tx;
        tr.alloc(sizeof(T) * (size_t)n);
        to.alloc(sizeof(T) * (size_t)n);
        tx.alloc(sizeof(T) * (size_t)n);
        SPB_CUDA(cudaMemsetAsync(tr.p, 0, tr.bytes, c->stream));
        SPB_CUDA(cudaMemsetAsync(to.p, 0, to.bytes, c->stream));
        if (op->wfwd.ok) wave_sweep<T>(op, op->wfwd, bufptr<T>(tr), bufptr<T>(to), bufptr<T>(tx));
        if (op->wbwd.ok) wave_sweep<T>(op, op->wbwd, bufptr<T>(tr), bufptr<T>(to), bufptr<T>(tx));
        SPB_CUDA(cudaStreamSynchronize(c->stream));
        if (getenv("SPB_GS_TIMING"))
          fprintf(stderr, "[gs_wave] launch shape: forward cluster %d, backward cluster %d (0 = plain CTAs)\n", op->wfwd.cluster, op->wbwd.cluster);
      }
    }
    // The global level schedule (cooperative grid-barrier kernel) is only the fallback: built now
    // when a wavefront schedule is unavailable, otherwise on first use (ensure_levels).
    if (!op->wfwd.ok || (mode == SPB_GS_SYMMETRIC && !op->wbwd.ok)) {
      build_levels(c, n, ip, cols, true, op->fwd);
      build_levels(c, n, ip, cols, false, op->bwd);
      op->levels_ready = true;
    }
  } catch (...) {
    delete op;
    throw;
  }
  return op;
}

// Fallback schedule on first use (a sweep the wavefront kernel does not take: the produced side is
// not the output vector).
template <typename T>
static void ensure_levels(GsOp<T>* M) {
  if (M->levels_ready) return;
  Ctx* c = M->ctx;
  CsrMat<T>* A = M->A;
  const int64_t n = A->n_local;
  std::vector<int64_t> ip(n + 1);
  std::vector<int> cols((size_t)A->nnz);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (A->ip64) {
    SPB_CUDA(cudaMemcpy(ip.data(), A->indptr.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost));
  } else {
    std::vector<int> ip32(n + 1);
    SPB_CUDA(cudaMemcpy(ip32.data(), A->indptr.p, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i <= n; ++i) ip[i] = ip32[i];
  }
  if (A->nnz) SPB_CUDA(cudaMemcpy(cols.data(), A->cols.p, sizeof(int) * A->nnz, cudaMemcpyDeviceToHost));
  build_levels(c, n, ip, cols, true, M->fwd);
  build_levels(c, n, ip, cols, false, M->bwd);
  M->levels_ready = true;
}

template <typename T, typename IP>
static void launch_sweep(GsOp<T>* M, const LevelSched& ls, const T* rhs, const T* lo, const T* hi, T* out, const T* xold) {
  Ctx* c = M->ctx;
  CsrMat<T>* A = M->A;
  if (A->n_local == 0) return;
  GsArgs<T, IP> a{bufptr<IP>(A->indptr), bufptr<int>(A->cols), bufptr<T>(A->vals), bufptr<T>(M->diag),
                  bufptr<int>(ls.level_ptr), bufptr<int>(ls.rows), (int)ls.nlevels, rhs, lo, hi, out,
                  bufptr<unsigned long long>(M->barrier), c->gate, c->gate_value,
                  (real_t<T>)M->omega != (real_t<T>)1 ? 1 : 0, (real_t<T>)M->omega, xold};
  int bps = 0;  // per call (per device): a process may hold contexts on several GPUs
  auto kern = gs_sweep_kernel<T, IP>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, 0);
  if (bps < 1) bps = 1;
  // no more CTAs than the widest level can use
  int64_t widest = 1;
  for (int64_t l = 0; l < ls.nlevels; ++l)
    widest = std::max<int64_t>(widest, ls.level_ptr_host[l + 1] - ls.level_ptr_host[l]);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)c->sm_count * std::min(bps, 4), ceil_div(widest, 256)));
  SPB_CUDA(cudaMemsetAsync(M->barrier.p, 0, sizeof(unsigned long long), c->stream));
  LaunchScope lsc(c, FAM_PRECOND);
  void* args[] = {(void*)&a};
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(256), args, 0, c->stream));
}

// xold: the values the relaxed update mixes with (null: zero); ignored by the plain sweep
template <typename T>
static void sweep(GsOp<T>* M, const LevelSched& ls, const T* rhs, const T* lo, const T* hi, T* out, const T* xold = nullptr) {
  // fast path: block-wavefront sweep (the produced side must be the vector being written; rhs and
  // the other triangle are consumed by its pre-pass, so out may alias them)
  const bool fwd = &ls == &M->fwd;
  WaveSched& ws = fwd ? M->wfwd : M->wbwd;
  const T* other = fwd ? hi : lo;
  if (ws.ok && (fwd ? lo : hi) == out) {
    wave_sweep<T>(M, ws, rhs, other, out);
    return;
  }
  ensure_levels(M);
  if (M->A->ip64)
    launch_sweep<T, int64_t>(M, ls, rhs, lo, hi, out, xold);
  else
    launch_sweep<T, int32_t>(M, ls, rhs, lo, hi, out, xold);
}

template <typename T>
void gs_apply(GsOp<T>* M, const T* in, T* out) {
  if (M->mode == SPB_GS_FORWARD) {
    sweep<T>(M, M->fwd, in, out, nullptr, out);
  } else {
    T* tmp = bufptr<T>(M->tmp);
    if (M->wfwd.ok && M->wbwd.ok && in != tmp && out != tmp) {
      // both sweeps on the wavefront path: the forward sweep hands its fold over the lower entries of every row to the
      // backward sweep, whose pre-pass would otherwise traverse the matrix once more to compute the very same sums
      if (!M->sig.p) M->sig.alloc(sizeof(T) * (size_t)std::max<int64_t>(M->A->n_local, 1));
      wave_sweep<T>(M, M->wfwd, in, nullptr, tmp, bufptr<T>(M->sig));
      wave_sweep<T>(M, M->wbwd, in, tmp, out, bufptr<T>(M->sig));
      return;
    }
    sweep<T>(M, M->fwd, in, tmp, nullptr, tmp);       // forward sweep from zero
    sweep<T>(M, M->bwd, in, tmp, out, out, tmp);      // rows n-1..0: lower cols = forward values (relaxed: mixed with them)
  }
}

template <typename T>
void gs_solver_sweep(GsOp<T>* M, const T* rhs, const T* x_old, T* x_new) {
  sweep<T>(M, M->fwd, rhs, x_new, x_old, x_new, x_old);
}

template <typename T>
void op_apply(spb_op* op, const T* in, T* out) {
  switch (op->kind) {
    case OP_CSR:
      static_cast<CsrMat<T>*>(op)->mul(in, out, EPI_NONE, nullptr, false);
      break;
    case OP_DIAG:
      diag_apply(static_cast<DiagOp<T>*>(op), in, out);
      break;
    case OP_GS:
      gs_apply(static_cast<GsOp<T>*>(op), in, out);
      break;
    default:
      SPB_FAIL(SPB_INVALID_ARG, "unknown operator kind");
  }
}

#define SPB_INST(T)                                                              \
  template DiagOp<T>* diag_from_host<T>(Ctx*, int, const void*, int64_t);        \
  template DiagOp<T>* diag_from_csr<T>(CsrMat<T>*);                              \
  template GsOp<T>* gs_create<T>(CsrMat<T>*, int, double);                       \
  template void csr_diagonal<T>(CsrMat<T>*, T*);                                 \
  template void diag_apply<T>(DiagOp<T>*, const T*, T*);                         \
  template void gs_apply<T>(GsOp<T>*, const T*, T*);                             \
  template void gs_solver_sweep<T>(GsOp<T>*, const T*, const T*, T*);            \
  template void op_apply<T>(spb_op*, const T*, T*);
SPB_INST(double)
SPB_INST(cplx)
SPB_INST(float)
SPB_INST(cplxf)

}  // namespace spb
