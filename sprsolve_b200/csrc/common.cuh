// common.cuh -- context, error plumbing, device buffers and launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/sprsolve_b200.h"
#include "scalar.cuh"

namespace spb {

void set_last_error(const std::string& msg);

struct SpbError {
  int status;
};

#define SPB_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,  \
               __LINE__, cudaGetErrorString(_e));                                             \
      ::spb::set_last_error(_b);                                                              \
      throw ::spb::SpbError{SPB_CUDA_ERROR};                                                  \
    }                                                                                         \
  } while (0)

#define SPB_FAIL(status, msg)             \
  do {                                    \
    ::spb::set_last_error(msg);           \
    throw ::spb::SpbError{(int)(status)}; \
  } while (0)

enum Family { FAM_SPMV = 0, FAM_VEC = 1, FAM_SCALAR = 2, FAM_PRECOND = 3, FAM_PACK = 4, FAM_COUNT = 5 };

struct Dist;  // dist.cu

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t comm_stream = nullptr;  // halo exchange runs here, overlapped with interior SpMV
  cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  int64_t fam_launches[FAM_COUNT] = {0, 0, 0, 0, 0};
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[FAM_COUNT];
  Dist* dist = nullptr;  // null => single GPU
  // Solver gate: while a device-resident solve is in flight, operator kernels run only if
  // *gate == gate_value (iterations queued past convergence become no-ops).
  const int* gate = nullptr;
  int gate_value = -1;
  int* dev_err = nullptr;  // device word set by kernels whose bounded spins timed out (1 halo, 2 Gauss-Seidel)
  void* pinned = nullptr;  // small pinned scratch for status / scalar read-backs
  size_t pinned_bytes = 0;

  int world() const;
  int rank() const;
};

// RAII device allocation.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) {
    o.p = nullptr;
    o.bytes = 0;
  }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      release();
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
      o.bytes = 0;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    SPB_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void ensure(size_t n) {
    if (n > bytes) alloc(n);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename U>
  U* as() const {
    return reinterpret_cast<U*>(p);
  }
};

// Launch accounting: every kernel of the library goes through launch_begin/launch_end so that
// bench.py can report gpu_launches and time one kernel family with CUDA events on the launching
// stream.
struct LaunchScope {
  Ctx* c;
  int fam;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  LaunchScope(Ctx* ctx, int family) : c(ctx), fam(family) {
    c->launches++;
    c->fam_launches[fam]++;
    if (c->profiling) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, c->stream);
    }
  }
  ~LaunchScope() {
    if (e0) {
      cudaEventRecord(e1, c->stream);
      c->prof_events[fam].push_back({e0, e1});
    }
  }
};

template <typename U>
inline U* bufptr(const DevBuf& b) {
  return reinterpret_cast<U*>(b.p);
}

inline void check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    char b[512];
    snprintf(b, sizeof(b), "kernel launch failed (%s): %s", what, cudaGetErrorString(e));
    set_last_error(b);
    throw SpbError{SPB_CUDA_ERROR};
  }
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- operator base (mirrors trait MatVecMul<T>, src/mat.rs:12-37) ---------------------------
enum OpKind { OP_CSR = 0, OP_DIAG = 1, OP_GS = 2 };

}  // namespace spb

struct spb_ctx : spb::Ctx {};

struct spb_op {
  spb::Ctx* ctx = nullptr;
  int kind = 0;
  int dtype = 0;
  int64_t n_global = 0;
  int64_t n_local = 0;
  int64_t row_begin = 0;
  virtual ~spb_op() {}
};
