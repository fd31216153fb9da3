// capi.cu -- the extern "C" boundary declared in include/sprsolve_b200.h.  Nothing throws across
// it: SpbError / std::exception are turned into spb_status codes + spb_last_error().
#include <cstring>
#include <new>
#include <stdexcept>

#include "csr.cuh"
#include "dist.cuh"
#include "ops.cuh"
#include "solver.cuh"
#include "vecops.cuh"

using namespace spb;

namespace spb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
}  // namespace spb

#define SPB_TRY try {
#define SPB_CATCH                                   \
  }                                                 \
  catch (const spb::SpbError& e) {                  \
    return e.status;                                \
  }                                                 \
  catch (const std::bad_alloc&) {                   \
    spb::set_last_error("host allocation failed");  \
    return SPB_CUDA_ERROR;                          \
  }                                                 \
  catch (const std::exception& e) {                 \
    spb::set_last_error(e.what());                  \
    return SPB_INVALID_ARG;                         \
  }

#define SPB_REQUIRE(cond, msg) \
  if (!(cond)) SPB_FAIL(SPB_INVALID_ARG, msg)

static void use_device(Ctx* c) { SPB_CUDA(cudaSetDevice(c->device)); }

// Runs `stmt` with `T` bound to the scalar type of a dtype code (s / d / c / z, src/mkl_mat.rs:68-71).
#define SPB_WITH_DTYPE(dt, ...)                        \
  switch (dt) {                                        \
    case SPB_F64: {                                    \
      using T = double;                                \
      __VA_ARGS__;                                     \
    } break;                                           \
    case SPB_C128: {                                   \
      using T = cplx;                                  \
      __VA_ARGS__;                                     \
    } break;                                           \
    case SPB_F32: {                                    \
      using T = float;                                 \
      __VA_ARGS__;                                     \
    } break;                                           \
    case SPB_C64: {                                    \
      using T = cplxf;                                 \
      __VA_ARGS__;                                     \
    } break;                                           \
    default:                                           \
      SPB_FAIL(SPB_INVALID_ARG, "bad dtype");          \
  }
static size_t dtype_size(int dt) {
  switch (dt) {
    case SPB_F64: return sizeof(double);
    case SPB_C128: return sizeof(cplx);
    case SPB_F32: return sizeof(float);
    case SPB_C64: return sizeof(cplxf);
    default: SPB_FAIL(SPB_INVALID_ARG, "bad dtype");
  }
}

// All functions below were declared extern "C" in include/sprsolve_b200.h and keep C linkage.

const char* spb_version(void) { return "sprsolve_b200 0.1.0 (sm_100a)"; }
const char* spb_last_error(void) { return g_last_error.c_str(); }

int spb_init(int device, spb_ctx** out) {
  SPB_TRY
  SPB_REQUIRE(out, "null out");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_last_error("no CUDA device: libsprsolve_b200 has no CPU fallback");
    return SPB_NO_DEVICE;
  }
  SPB_REQUIRE(device >= 0 && device < ndev, "bad device index");
  SPB_CUDA(cudaSetDevice(device));
  auto* c = new spb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  SPB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  SPB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  SPB_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
  SPB_CUDA(cudaEventCreateWithFlags(&c->ev_pack, cudaEventDisableTiming));
  SPB_CUDA(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
  SPB_CUDA(cudaMalloc(&c->dev_err, 4 * sizeof(int)));
  SPB_CUDA(cudaMemset(c->dev_err, 0, 4 * sizeof(int)));
  *out = c;
  return SPB_OK;
  SPB_CATCH
}

int spb_finalize(spb_ctx* c) {
  SPB_TRY
  if (!c) return SPB_OK;
  use_device(c);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->comm_stream);
  for (int f = 0; f < FAM_COUNT; ++f)
    for (auto& p : c->prof_events[f]) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
  if (c->dist) {
    if (c->dist->comm_halo && c->dist->comm_halo != c->dist->comm) nccl().CommDestroy(c->dist->comm_halo);
    if (c->dist->comm) nccl().CommDestroy(c->dist->comm);
    window_destroy(c->dist->scal);
    delete c->dist;
  }
  if (c->own_stream) cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->comm_stream);
  cudaEventDestroy(c->ev_pack);
  cudaEventDestroy(c->ev_halo);
  if (c->dev_err) cudaFree(c->dev_err);
  delete c;
  return SPB_OK;
  SPB_CATCH
}

int spb_set_stream(spb_ctx* c, void* stream) {
  SPB_TRY
  SPB_REQUIRE(c, "null ctx");
  use_device(c);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)stream;
  c->own_stream = false;
  return SPB_OK;
  SPB_CATCH
}

int spb_synchronize(spb_ctx* c) {
  SPB_TRY
  SPB_REQUIRE(c, "null ctx");
  use_device(c);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->comm_stream));
  device_check(c);  // a bounded device-side spin that timed out (dead peer, stalled sweep) surfaces here
  return SPB_OK;
  SPB_CATCH
}

int64_t spb_launch_count(spb_ctx* c) { return c ? c->launches : 0; }

int spb_profile_enable(spb_ctx* c, int on) {
  SPB_TRY
  SPB_REQUIRE(c, "null ctx");
  c->profiling = on != 0;
  return SPB_OK;
  SPB_CATCH
}

int spb_profile_read(spb_ctx* c, int family, int64_t* launches, double* ms) {
  SPB_TRY
  SPB_REQUIRE(c && family >= 0 && family < FAM_COUNT, "bad family");
  use_device(c);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (auto& p : c->prof_events[family]) {
    float t = 0.f;
    SPB_CUDA(cudaEventSynchronize(p.second));
    SPB_CUDA(cudaEventElapsedTime(&t, p.first, p.second));
    tot += t;
  }
  if (launches) *launches = (int64_t)c->prof_events[family].size();
  if (ms) *ms = tot;
  return SPB_OK;
  SPB_CATCH
}

int spb_profile_reset(spb_ctx* c) {
  SPB_TRY
  SPB_REQUIRE(c, "null ctx");
  use_device(c);
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  for (int f = 0; f < FAM_COUNT; ++f) {
    for (auto& p : c->prof_events[f]) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
    c->prof_events[f].clear();
  }
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- communicator
int spb_comm_unique_id(void* id128) {
  SPB_TRY
  SPB_REQUIRE(id128, "null id");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  SPB_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return SPB_OK;
  SPB_CATCH
}

int spb_stencil_partition(int kind, int64_t nx, int64_t ny, int64_t nz, int world, int rank,
                          int64_t* row_begin, int64_t* row_end) {
  SPB_TRY
  SPB_REQUIRE(row_begin && row_end && world >= 1 && rank >= 0 && rank < world, "bad partition arguments");
  SPB_REQUIRE(kind >= 0 && kind <= SPB_STENCIL_CONVDIFF27 && nx >= 1 && ny >= 1 && nz >= 1, "bad grid");
  stencil_partition(kind, nx, ny, kind == SPB_STENCIL_DIRICHLET2D ? 1 : nz, world, rank, row_begin, row_end);
  return SPB_OK;
  SPB_CATCH
}

int spb_comm_init(spb_ctx* c, int world, int rank, const void* id128) {
  SPB_TRY
  SPB_REQUIRE(c && id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  SPB_REQUIRE(!c->dist, "communicator already initialised");
  use_device(c);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  auto* d = new Dist();
  d->world = world;
  d->rank = rank;
  try {
    SPB_NCCL(nccl().CommInitRank(&d->comm, world, id, rank));
    if (nccl().CommSplit) {
      SPB_NCCL(nccl().CommSplit(d->comm, 0, rank, &d->comm_halo, nullptr));
    } else {
      d->comm_halo = d->comm;
    }
  } catch (...) {
    delete d;
    throw;
  }
  c->dist = d;
  // Data-path transport: stores into peer-mapped windows over NVLink (default) or NCCL
  // (SPB_COMM=nccl; also the automatic choice when some peer cannot be mapped through CUDA IPC).
  const char* tr = getenv("SPB_COMM");
  if (world > 1 && !(tr && strcmp(tr, "nccl") == 0)) {
    d->scal = window_create(c, sizeof(ScalWin));
    d->peer = d->scal != nullptr;
  }
  return SPB_OK;
  SPB_CATCH
}

int spb_comm_info(spb_ctx* c, int* world, int* rank) {
  SPB_TRY
  SPB_REQUIRE(c, "null ctx");
  if (world) *world = c->world();
  if (rank) *rank = c->rank();
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- matrices
int spb_csr_create(spb_ctx* c, int dtype, int64_t nrows, int64_t ncols, int64_t row_begin,
                   int64_t row_end, const void* indptr, int indptr_bits, const int32_t* indices,
                   const void* values, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out && indptr, "null argument");
  *out = nullptr;
  use_device(c);
  if (nrows != ncols) {  // assert_eq!(ncol, nrow), src/mkl_mat.rs:37
    set_last_error("Not a square matrix");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  if (!c->dist) SPB_REQUIRE(row_begin == 0 && row_end == nrows, "row range needs a communicator");
  SPB_WITH_DTYPE(dtype, *out = csr_from_host<T>(c, nrows, row_begin, row_end, indptr, indptr_bits, indices, values));
  return SPB_OK;
  SPB_CATCH
}

int spb_csr_create_stencil(spb_ctx* c, int kind, int dtype, int64_t nx, int64_t ny, int64_t nz,
                           const double* params, int nparams, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out, "null argument");
  *out = nullptr;
  use_device(c);
  SPB_WITH_DTYPE(dtype, *out = csr_from_stencil<T>(c, kind, nx, ny, nz, params, nparams));
  return SPB_OK;
  SPB_CATCH
}

int spb_csc_create(spb_ctx* c, int dtype, int64_t nrows, int64_t ncols, const void* indptr, int indptr_bits,
                   const int32_t* row_indices, const void* values, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out && indptr, "null argument");
  *out = nullptr;
  use_device(c);
  if (nrows != ncols) {
    set_last_error("Not a square matrix");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  SPB_WITH_DTYPE(dtype, *out = csr_from_csc<T>(c, nrows, indptr, indptr_bits, row_indices, values));
  return SPB_OK;
  SPB_CATCH
}

int spb_csr_create_from_triplets(spb_ctx* c, int dtype, int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rows,
                                 const int32_t* cols, const void* values, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out && (nnz == 0 || (rows && cols && values)), "null argument");
  *out = nullptr;
  use_device(c);
  if (nrows != ncols) {
    set_last_error("Not a square matrix");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  SPB_WITH_DTYPE(dtype, *out = csr_from_triplets<T>(c, nrows, nnz, rows, cols, values));
  return SPB_OK;
  SPB_CATCH
}

int spb_csr_read_matrix_market(spb_ctx* c, int dtype, const char* path, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out && path, "null argument");
  *out = nullptr;
  use_device(c);
  SPB_WITH_DTYPE(dtype, *out = csr_from_matrix_market<T>(c, path));
  return SPB_OK;
  SPB_CATCH
}

static int rehint(spb_op* m) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR, "not a CSR matrix");
  use_device(m->ctx);
  SPB_WITH_DTYPE(m->dtype, static_cast<CsrMat<T>*>(m)->autotune());
  return SPB_OK;
  SPB_CATCH
}
int spb_csr_mv_hint(spb_op* m, int) { return rehint(m); }
int spb_csr_mv_and_dotmv_hint(spb_op* m, int) { return rehint(m); }

int spb_op_size(spb_op* op, int64_t* n_global, int64_t* n_local, int64_t* row_begin) {
  SPB_TRY
  SPB_REQUIRE(op, "null op");
  if (n_global) *n_global = op->n_global;
  if (n_local) *n_local = op->n_local;
  if (row_begin) *row_begin = op->row_begin;
  return SPB_OK;
  SPB_CATCH
}

int spb_csr_nnz(spb_op* m, int64_t* nnz) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR && nnz, "not a CSR matrix");
  SPB_WITH_DTYPE(m->dtype, *nnz = static_cast<CsrMat<T>*>(m)->nnz);
  return SPB_OK;
  SPB_CATCH
}

template <typename T>
static void csr_download(CsrMat<T>* m, int64_t* indptr64, int32_t* indices, void* values) {
  Ctx* c = m->ctx;
  const int64_t n = m->n_local;
  if (indptr64) {
    if (m->ip64) {
      SPB_CUDA(cudaMemcpyAsync(indptr64, m->indptr.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
      std::vector<int32_t> t(n + 1);
      SPB_CUDA(cudaMemcpyAsync(t.data(), m->indptr.p, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, c->stream));
      SPB_CUDA(cudaStreamSynchronize(c->stream));
      for (int64_t i = 0; i <= n; ++i) indptr64[i] = t[i];
    }
  }
  if (indices && m->nnz) {
    SPB_CUDA(cudaMemcpyAsync(indices, m->cols.p, sizeof(int32_t) * m->nnz, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->dist) {  // local + halo ids back to global ids
      for (int64_t k = 0; k < m->nnz; ++k) {
        const int32_t col = indices[k];
        indices[k] = col < m->n_local ? (int32_t)(col + m->row_begin) : m->halo_cols_global[col - m->n_local];
      }
    }
  }
  if (values && m->nnz) {
    SPB_CUDA(cudaMemcpyAsync(values, m->vals.p, sizeof(T) * m->nnz, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
  }
}

namespace {
template <typename T>
void csr_plan_info_impl(spb_op* op, int64_t* info) {
  auto* m = static_cast<CsrMat<T>*>(op);
  const int64_t ipb = m->ip64 ? 8 : 4, vb = (int64_t)sizeof(T);
  // bytes one mul_vec streams from HBM by design: values + x once + y once + either column indices and row pointers
  // (plain CSR) or one 32-bit word per row (dictionary: 16-bit pattern id + the low 16 bits of the row pointer)
  const int64_t stream = m->dict_on ? m->nnz * vb + 4 * (m->n_local + 1) + 2 * m->n_local * vb
                                    : m->nnz * (vb + 4) + (m->n_local + 1) * ipb + 2 * m->n_local * vb;
  // (info[0]: 0 plain CSR stream, 1 dictionary, 3 dictionary + x window in shared memory)
  const int64_t v[8] = {(m->dict_on ? 1 : 0) | (m->dict_on && m->xwin_on ? 2 : 0), m->dict_u, m->dict_w, m->plan_ct, m->plan_stages, m->plan_tile, m->plan_bps, stream};
  for (int i = 0; i < 8; ++i) info[i] = v[i];
}
}  // namespace

int spb_csr_plan_info(spb_op* m, int64_t info[8]) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR && info, "not a CSR matrix");
  SPB_WITH_DTYPE(m->dtype, csr_plan_info_impl<T>(m, info));
  return SPB_OK;
  SPB_CATCH
}

int spb_csr_download(spb_op* m, int64_t* indptr64, int32_t* indices, void* values) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR, "not a CSR matrix");
  use_device(m->ctx);
  SPB_WITH_DTYPE(m->dtype, csr_download(static_cast<CsrMat<T>*>(m), indptr64, indices, values));
  return SPB_OK;
  SPB_CATCH
}

template <typename T>
static void csr_diag_to_host(CsrMat<T>* m, void* host) {
  DevBuf d;
  d.alloc(sizeof(T) * (size_t)std::max<int64_t>(m->n_local, 1));
  csr_diagonal<T>(m, bufptr<T>(d));
  SPB_CUDA(cudaMemcpyAsync(host, d.p, sizeof(T) * m->n_local, cudaMemcpyDeviceToHost, m->ctx->stream));
  SPB_CUDA(cudaStreamSynchronize(m->ctx->stream));
}

int spb_csr_diagonal(spb_op* m, void* diag_host) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR && diag_host, "not a CSR matrix");
  use_device(m->ctx);
  SPB_WITH_DTYPE(m->dtype, csr_diag_to_host(static_cast<CsrMat<T>*>(m), diag_host));
  return SPB_OK;
  SPB_CATCH
}

int spb_op_destroy(spb_op* op) {
  SPB_TRY
  if (!op) return SPB_OK;
  use_device(op->ctx);
  cudaStreamSynchronize(op->ctx->stream);
  delete op;
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- trait methods
template <typename T>
static void op_mul_dev(spb_op* op, const T* in, T* out, bool with_dot, double* dot_out) {
  Ctx* c = op->ctx;
  if (!with_dot) {
    op_apply<T>(op, in, out);
    return;
  }
  if (op->kind != OP_CSR) {
    set_last_error("mul_vec_dot is unimplemented for this operator (src/precond.rs:55-62)");
    throw SpbError{SPB_UNIMPLEMENTED};
  }
  auto* m = static_cast<CsrMat<T>*>(op);
  m->mul(in, out, EPI_DOT_WY, in, false);  // conj(v_in) . v_out fused (mkl_sparse_?_dotmv analogue)
  m->finalize_epilogue(true);
  scal2 h[2];
  SPB_CUDA(cudaMemcpyAsync(h, m->red.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  dot_out[0] = h[0].re;
  dot_out[1] = h[0].im;
}

template <typename T>
static void op_mul_host(spb_op* op, const void* v_in, void* v_out, bool with_dot, double* dot_out) {
  Ctx* c = op->ctx;
  const int64_t n = op->n_local;
  DevBuf din, dout;
  din.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
  dout.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
  if (n) SPB_CUDA(cudaMemcpyAsync(din.p, v_in, sizeof(T) * n, cudaMemcpyHostToDevice, c->stream));
  op_mul_dev<T>(op, bufptr<T>(din), bufptr<T>(dout), with_dot, dot_out);
  if (n) SPB_CUDA(cudaMemcpyAsync(v_out, dout.p, sizeof(T) * n, cudaMemcpyDeviceToHost, c->stream));
  device_check(c);  // synchronises; a timed-out halo / sweep spin must not hand back stale data as SPB_OK
}

static int op_mul_checked(spb_op* op, const void* v_in, int64_t n_in, void* v_out, int64_t n_out,
                          bool with_dot, double* dot_out) {
  SPB_TRY
  SPB_REQUIRE(op && (v_in || n_in == 0) && (v_out || n_out == 0), "null argument");
  use_device(op->ctx);
  if (op->n_local != n_in || n_in != n_out) {  // src/mat.rs:50-52, mkl_mat.rs:154-156, precond.rs:39-41
    set_last_error("Dimension mismatch");
    return SPB_DIM_MISMATCH;
  }
  SPB_WITH_DTYPE(op->dtype, op_mul_host<T>(op, v_in, v_out, with_dot, dot_out));
  return SPB_OK;
  SPB_CATCH
}

int spb_op_mul_vec(spb_op* op, const void* v_in, int64_t n_in, void* v_out, int64_t n_out) {
  return op_mul_checked(op, v_in, n_in, v_out, n_out, false, nullptr);
}
int spb_op_mul_vec_dot(spb_op* op, const void* v_in, int64_t n_in, void* v_out, int64_t n_out, double out[2]) {
  if (!out) return SPB_INVALID_ARG;
  return op_mul_checked(op, v_in, n_in, v_out, n_out, true, out);
}
int spb_op_mul_vec_dev(spb_op* op, const void* d_in, void* d_out) {
  SPB_TRY
  SPB_REQUIRE(op && d_in && d_out, "null argument");
  use_device(op->ctx);
  SPB_WITH_DTYPE(op->dtype, op_mul_dev<T>(op, (const T*)d_in, (T*)d_out, false, nullptr));
  return SPB_OK;
  SPB_CATCH
}
int spb_op_mul_vec_dot_dev(spb_op* op, const void* d_in, void* d_out, double out[2]) {
  SPB_TRY
  SPB_REQUIRE(op && d_in && d_out && out, "null argument");
  use_device(op->ctx);
  SPB_WITH_DTYPE(op->dtype, op_mul_dev<T>(op, (const T*)d_in, (T*)d_out, true, out));
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- preconditioners
int spb_diag_precond_create(spb_ctx* c, int dtype, int diag_dtype, const void* diag, int64_t n, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(c && out && (diag || n == 0) && n >= 0, "bad argument");
  *out = nullptr;
  use_device(c);
  // V = T, or V = T::Real for a complex system (DiagPrecond<Complex64, f64> / <Complex32, f32>)
  SPB_WITH_DTYPE(dtype, {
    SPB_REQUIRE(diag_dtype == dtype || diag_dtype == ScalarTraits<real_t<T>>::dtype,
                "the diagonal must have the system's scalar type or its real type");
    *out = diag_from_host<T>(c, diag_dtype, diag, n);
  });
  return SPB_OK;
  SPB_CATCH
}

int spb_diag_precond_from_csr(spb_op* m, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(m && m->kind == OP_CSR && out, "not a CSR matrix");
  use_device(m->ctx);
  SPB_WITH_DTYPE(m->dtype, *out = diag_from_csr(static_cast<CsrMat<T>*>(m)));
  return SPB_OK;
  SPB_CATCH
}

int spb_gs_precond_create(spb_op* m, int mode, spb_op** out) { return spb_gs_precond_create_relaxed(m, mode, 1.0, out); }

int spb_gs_precond_create_relaxed(spb_op* m, int mode, double omega, spb_op** out) {
  SPB_TRY
  SPB_REQUIRE(m && out, "null argument");
  *out = nullptr;
  if (m->kind != OP_CSR) {
    set_last_error("Not in CSR format");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  use_device(m->ctx);
  spb_op* op = nullptr;
  int64_t bad = -1;
  SPB_WITH_DTYPE(m->dtype, {
    auto* g = gs_create(static_cast<CsrMat<T>*>(m), mode, omega);
    bad = g->bad_row;
    op = g;
  });
  if (bad >= 0) {
    delete op;
    char b[128];
    snprintf(b, sizeof(b), "Matrix has zero diagonal element at %lld", (long long)bad);
    set_last_error(b);
    return SPB_ZERO_DIAGONAL;
  }
  *out = op;
  return SPB_OK;
  SPB_CATCH
}

int spb_gs_levels(spb_op* gs, int64_t* nf, int64_t* nb) {
  SPB_TRY
  SPB_REQUIRE(gs && gs->kind == OP_GS, "not a Gauss-Seidel operator");
  SPB_WITH_DTYPE(gs->dtype, {
    auto* g = static_cast<GsOp<T>*>(gs);
    if (nf) *nf = g->wfwd.ok ? g->wfwd.global_levels : g->fwd.nlevels;
    if (nb) *nb = g->wbwd.ok ? g->wbwd.global_levels : g->bwd.nlevels;
  });
  return SPB_OK;
  SPB_CATCH
}

namespace {
template <typename T>
void gs_schedule_info_impl(spb_op* gs, int64_t* info, int64_t* stats, int64_t cap) {
  auto* g = static_cast<GsOp<T>*>(gs);
  Ctx* c = g->ctx;
  int flags[4] = {0, 0, 0, 0}, flagsb[4] = {0, 0, 0, 0};
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  if (g->wfwd.ok) SPB_CUDA(cudaMemcpy(flags, g->wfwd.ticket.p, sizeof(flags), cudaMemcpyDeviceToHost));
  if (g->wbwd.ok) SPB_CUDA(cudaMemcpy(flagsb, g->wbwd.ticket.p, sizeof(flagsb), cudaMemcpyDeviceToHost));
  const int64_t v[16] = {g->wfwd.ok, g->wfwd.block_rows, g->wfwd.nblocks, g->wfwd.nchunks, g->wfwd.local_levels_max,
                         g->wfwd.stages, (int64_t)g->wfwd.smem_bytes, g->wbwd.ok, g->wbwd.nchunks, g->wbwd.local_levels_max,
                         flags[1] | flagsb[1], g->wfwd.rhs_slots, g->wfwd.aux_slots, (int64_t)g->wfwd.stat.bytes, 0, 0};
  if (info) for (int i = 0; i < 16; ++i) info[i] = v[i];
  if (stats && g->wave_stats.p) {
    const int64_t m = std::min<int64_t>(cap, 4 * (int64_t)g->wfwd.nblocks);
    SPB_CUDA(cudaMemcpy(stats, g->wave_stats.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
  }
}
}  // namespace

int spb_gs_schedule_info(spb_op* gs, int64_t info[16], int64_t* stats, int64_t stats_cap) {
  SPB_TRY
  SPB_REQUIRE(gs && gs->kind == OP_GS, "not a Gauss-Seidel operator");
  SPB_WITH_DTYPE(gs->dtype, gs_schedule_info_impl<T>(gs, info, stats, stats_cap));
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- vecalg on host slices
namespace {
template <typename T>
struct HostVec {
  DevBuf d;
  Ctx* c;
  int64_t n;
  HostVec(Ctx* ctx, const void* h, int64_t n_) : c(ctx), n(n_) {
    d.alloc(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
    if (h && n) SPB_CUDA(cudaMemcpyAsync(d.p, h, sizeof(T) * n, cudaMemcpyHostToDevice, c->stream));
  }
  T* p() { return bufptr<T>(d); }
  void back(void* h) {
    if (n) SPB_CUDA(cudaMemcpyAsync(h, d.p, sizeof(T) * n, cudaMemcpyDeviceToHost, c->stream));
    SPB_CUDA(cudaStreamSynchronize(c->stream));
  }
};

template <typename T>
void vec_reduce_host(Ctx* c, int kind, int64_t n, const void* x, const void* y, double out[2]) {
  HostVec<T> dx(c, x, n), dy(c, y ? y : x, n);
  DevBuf parts, red;
  parts.alloc(sizeof(Acc<T>) * 2 * (size_t)(vec_max_grid(c) + 1));
  red.alloc(sizeof(scal2) * 2);
  vec_reduce<T>(c, kind, n, dx.p(), dy.p(), bufptr<Acc<T>>(parts), bufptr<scal2>(red));
  scal2 h[2];
  SPB_CUDA(cudaMemcpyAsync(h, red.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  out[0] = h[0].re;
  out[1] = h[0].im;
}
template <typename T>
T scalar_of(const double a[2]);
template <>
double scalar_of<double>(const double a[2]) {
  return a[0];
}
template <>
cplx scalar_of<cplx>(const double a[2]) {
  return cplx{a[0], a[1]};
}
template <>
float scalar_of<float>(const double a[2]) {
  return (float)a[0];
}
template <>
cplxf scalar_of<cplxf>(const double a[2]) {
  return cplxf{(float)a[0], (float)a[1]};
}
}  // namespace

namespace {
template <typename T>
void scale_host(Ctx* c, int64_t n, const double a[2], void* x) {
  HostVec<T> dx(c, x, n);
  vec_scale<T>(c, n, scalar_of<T>(a), dx.p());
  dx.back(x);
}
template <typename T>
void rscale_host(Ctx* c, int64_t n, double a, void* x) {
  HostVec<T> dx(c, x, n);
  vec_rscale<T>(c, n, (real_t<T>)a, dx.p());
  dx.back(x);
}
template <typename T>
void conj_host(Ctx* c, int64_t n, const void* x, void* out) {
  HostVec<T> dx(c, x, n);
  HostVec<T> dy(c, nullptr, n);
  vec_conj<T>(c, n, dx.p(), dy.p());
  dy.back(out);
}
template <typename T>
void axpy_host(Ctx* c, int64_t n, const double a[2], const void* x, void* y) {
  HostVec<T> dx(c, x, n);
  HostVec<T> dy(c, y, n);
  vec_axpy<T>(c, n, scalar_of<T>(a), dx.p(), dy.p());
  dy.back(y);
}
template <typename T>
void axpby_host(Ctx* c, int64_t n, const double a[2], const void* x, const double b[2], void* y) {
  HostVec<T> dx(c, x, n);
  HostVec<T> dy(c, y, n);
  vec_axpby<T>(c, n, scalar_of<T>(a), dx.p(), scalar_of<T>(b), dy.p());
  dy.back(y);
}
}  // namespace

#define SPB_DISPATCH(dtype, fn, ...) SPB_WITH_DTYPE(dtype, fn<T>(__VA_ARGS__))

int spb_vec_dot(spb_ctx* c, int dtype, int64_t n, const void* x, const void* y, double out[2]) {
  SPB_TRY
  SPB_REQUIRE(c && out && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, vec_reduce_host, c, 0, n, x, y, out);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_conj_dot(spb_ctx* c, int dtype, int64_t n, const void* x, const void* y, double out[2]) {
  SPB_TRY
  SPB_REQUIRE(c && out && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, vec_reduce_host, c, 1, n, x, y, out);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_norm2(spb_ctx* c, int dtype, int64_t n, const void* x, double* out) {
  SPB_TRY
  SPB_REQUIRE(c && out && n >= 0, "bad argument");
  use_device(c);
  double r[2];
  SPB_DISPATCH(dtype, vec_reduce_host, c, 2, n, x, nullptr, r);
  // src/vecalg.rs:603-604: the sum and the square root are T::Real
  *out = (dtype == SPB_F32 || dtype == SPB_C64) ? (double)sqrtf((float)r[0]) : sqrt(r[0]);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_scale(spb_ctx* c, int dtype, int64_t n, const double a[2], void* x) {
  SPB_TRY
  SPB_REQUIRE(c && a && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, scale_host, c, n, a, x);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_rscale(spb_ctx* c, int dtype, int64_t n, double a, void* x) {
  SPB_TRY
  SPB_REQUIRE(c && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, rscale_host, c, n, a, x);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_conj(spb_ctx* c, int dtype, int64_t n, const void* x, void* out) {
  SPB_TRY
  SPB_REQUIRE(c && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, conj_host, c, n, x, out);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_axpy(spb_ctx* c, int dtype, int64_t n, const double a[2], const void* x, void* y) {
  SPB_TRY
  SPB_REQUIRE(c && a && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, axpy_host, c, n, a, x, y);
  return SPB_OK;
  SPB_CATCH
}
int spb_vec_axpby(spb_ctx* c, int dtype, int64_t n, const double a[2], const void* x, const double b[2], void* y) {
  SPB_TRY
  SPB_REQUIRE(c && a && b && n >= 0, "bad argument");
  use_device(c);
  SPB_DISPATCH(dtype, axpby_host, c, n, a, x, b, y);
  return SPB_OK;
  SPB_CATCH
}

// ---------------------------------------------------------------- solvers
static int make_solver(spb_op* A, int64_t size, int which, spb_solver** out, double omega = 1.0) {
  SPB_TRY
  SPB_REQUIRE(A && out && size >= 0, "bad argument");
  *out = nullptr;
  use_device(A->ctx);
  if (which == 3)
    *out = make_gauss_seidel(A, omega);
  else {
    SPB_REQUIRE(A->kind == OP_CSR, "solver operator must be a CSR matrix");
    *out = which == 0 ? make_bicgstab(A, size) : make_minres(A, size, which == 2);
  }
  SPB_CUDA(cudaStreamSynchronize(A->ctx->stream));
  return SPB_OK;
  SPB_CATCH
}
int spb_bicgstab_create(spb_op* A, int64_t size, spb_solver** out) { return make_solver(A, size, 0, out); }
int spb_minres_create(spb_op* A, int64_t size, spb_solver** out) { return make_solver(A, size, 1, out); }
int spb_csminres_create(spb_op* A, int64_t size, spb_solver** out) { return make_solver(A, size, 2, out); }
int spb_gauss_seidel_create(spb_op* A, spb_solver** out) { return make_solver(A, A ? A->n_local : 0, 3, out); }
int spb_gauss_seidel_create_relaxed(spb_op* A, double omega, spb_solver** out) {
  return make_solver(A, A ? A->n_local : 0, 3, out, omega);
}

int spb_solver_solve_dev(spb_solver* s, spb_op* precond, const void* d_rhs, void* d_x, int64_t max_iter,
                         double tol, int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                         int64_t* hist_len) {
  SPB_TRY
  SPB_REQUIRE(s && d_rhs && d_x && iters && resid && max_iter >= 0, "bad argument");
  use_device(s->ctx);
  *iters = 0;
  *resid = 0.0;
  if (hist_len) *hist_len = 0;
  if (precond && precond->dtype != s->dtype) SPB_FAIL(SPB_INVALID_ARG, "preconditioner dtype mismatch");
  const int rc = s->solve_dev(precond, d_rhs, d_x, max_iter, tol, iters, resid, hist, hist_cap, hist_len);
  device_check(s->ctx);
  return rc;
  SPB_CATCH
}

int spb_solver_solve(spb_solver* s, spb_op* precond, const void* rhs, int64_t n_rhs, void* x, int64_t n_x,
                     int64_t max_iter, double tol, int64_t* iters, double* resid, double* hist,
                     int64_t hist_cap, int64_t* hist_len) {
  SPB_TRY
  SPB_REQUIRE(s && iters && resid && max_iter >= 0, "bad argument");
  use_device(s->ctx);
  *iters = 0;
  *resid = 0.0;
  if (hist_len) *hist_len = 0;
  // dimension checks of the reference (src/bicg_stab.rs:44-53, src/gauss_seidel.rs:41-50)
  if (n_rhs != s->size) {
    set_last_error("Input vec dimension doesn't match the matrix size");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  if (n_rhs != n_x) {
    set_last_error("Input and output vec dimension do not match");
    return SPB_INCOMPATIBLE_FORMAT;
  }
  SPB_REQUIRE((rhs && x) || n_rhs == 0, "null vector");
  if (precond && precond->dtype != s->dtype) SPB_FAIL(SPB_INVALID_ARG, "preconditioner dtype mismatch");
  Ctx* c = s->ctx;
  const size_t esz = dtype_size(s->dtype);
  const size_t bytes = esz * (size_t)std::max<int64_t>(n_rhs, 1);
  s->stage_rhs.ensure(bytes);
  s->stage_x.ensure(bytes);
  if (n_rhs) {
    SPB_CUDA(cudaMemcpyAsync(s->stage_rhs.p, rhs, esz * n_rhs, cudaMemcpyHostToDevice, c->stream));
    SPB_CUDA(cudaMemcpyAsync(s->stage_x.p, x, esz * n_rhs, cudaMemcpyHostToDevice, c->stream));
  }
  const int rc = s->solve_dev(precond, s->stage_rhs.p, s->stage_x.p, max_iter, tol, iters, resid, hist, hist_cap, hist_len);
  device_check(c);
  if (n_rhs) SPB_CUDA(cudaMemcpyAsync(x, s->stage_x.p, esz * n_rhs, cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaStreamSynchronize(c->stream));
  return rc;
  SPB_CATCH
}

int spb_solver_set_poll_interval(spb_solver* s, int iters) {
  SPB_TRY
  SPB_REQUIRE(s && iters >= 1, "bad argument");
  s->poll = iters;
  return SPB_OK;
  SPB_CATCH
}

int spb_solver_destroy(spb_solver* s) {
  SPB_TRY
  if (!s) return SPB_OK;
  use_device(s->ctx);
  cudaStreamSynchronize(s->ctx->stream);
  delete s;
  return SPB_OK;
  SPB_CATCH
}

