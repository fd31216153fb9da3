// solver.cuh -- shared base of the device-resident solvers.
//
// The reference's solvers (src/bicg_stab.rs, src/minres.rs, src/cs_minres.rs,
// src/gauss_seidel.rs) run their loop on the host and call the operator on host slices every
// iteration.  Here the whole loop runs on the device: Krylov scalars, the iteration counter and
// the status word live in a device-side state struct, every kernel starts by checking the status
// word (so iterations queued past convergence are no-ops), and the host only polls that word
// every `poll` iterations through pinned memory -- lagging one chunk behind, so the launch queue
// never drains.
#pragma once
#include "csr.cuh"
#include "dist.cuh"
#include "ops.cuh"
#include "vecops.cuh"

namespace spb {

// device-side status word
enum DevStatus {
  DS_RUNNING = -1,
  DS_OK = SPB_OK,
  DS_BREAKDOWN = SPB_BREAKDOWN,
  DS_INVALID_PRECOND = SPB_INVALID_PRECOND,
  DS_ZERO_RHS = 50,      // ||b|| <= eps: x := 0, Ok((0, ||b||))   (src/bicg_stab.rs:55-60)
  DS_NEED_RESTART = 51   // BiCGStab rho restart requested (src/bicg_stab.rs:131-145)
};

enum PcMode { PCM_NONE = 0, PCM_JACOBI = 1, PCM_JACOBI_REAL = 2, PCM_GENERIC = 3 };

// What every solver state begins with (read back by the host when polling).
struct StateHead {
  int status;
  int pad;
  long long its;        // BiCGStab: index of the iteration being executed
  long long res_iters;  // value returned in Ok((iters, resid)) / BreakDown(its)
  double res_resid;
  long long hist_len;
};

}  // namespace spb

struct spb_solver {
  spb::Ctx* ctx = nullptr;
  int kind = 0;  // 0 bicgstab, 1 minres, 2 csminres, 3 gauss-seidel
  int dtype = 0;
  spb_op* A = nullptr;
  int64_t size = 0;  // the `size` handed to ::new (dimension checks, src/bicg_stab.rs:44)
  int poll = 16;
  spb::DevBuf stage_rhs, stage_x;  // host-slice entry point staging
  virtual ~spb_solver() {}
  // device pointers, local length == size
  virtual int solve_dev(spb_op* M, const void* d_rhs, void* d_x, int64_t max_iter, double tol,
                        int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                        int64_t* hist_len) = 0;
};

namespace spb {

template <typename T>
inline PcMode pc_mode_of(spb_op* M) {
  if (!M) return PCM_NONE;
  if (M->kind == OP_DIAG) return static_cast<DiagOp<T>*>(M)->real_diag ? PCM_JACOBI_REAL : PCM_JACOBI;
  return PCM_GENERIC;
}

spb_solver* make_bicgstab(spb_op* A, int64_t size);
spb_solver* make_minres(spb_op* A, int64_t size, bool cs);
spb_solver* make_gauss_seidel(spb_op* A, double omega = 1.0);

// Lagged status polling through pinned memory.
struct Poller {
  Ctx* c;
  StateHead* pinned[2];
  cudaEvent_t ev[2];
  bool pending[2] = {false, false};
  int cur = 0;
  explicit Poller(Ctx* ctx);
  ~Poller();
  // queue an async copy of the state head
  void post(const void* d_state);
  // wait for the OLDER outstanding copy (if any) and return it; returns false if none pending
  bool wait_oldest(StateHead* out);
  // wait for all outstanding copies, return the newest
  bool drain(StateHead* out);
};

}  // namespace spb
