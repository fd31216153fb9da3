// ops.cuh -- preconditioner operators implementing MatVecMul<T>: DiagPrecond (src/precond.rs)
// and the level-scheduled Gauss-Seidel sweep operator (sweep body: src/gauss_seidel.rs:111-125).
#pragma once
#include "csr.cuh"

namespace spb {

template <typename T>
struct DiagOp : spb_op {
  DevBuf dinv;        // T or double (real_diag) [n]: 1/diag, src/precond.rs:20-24
  bool real_diag = false;
};

// Level schedule of one triangular dependency pattern.
struct LevelSched {
  int64_t nlevels = 0;
  DevBuf level_ptr;   // int32 [nlevels+1]
  DevBuf rows;        // int32 [n]: rows sorted by level (ascending row id inside a level)
  std::vector<int> level_ptr_host;
};

template <typename T>
struct GsOp : spb_op {
  CsrMat<T>* A = nullptr;
  int mode = SPB_GS_FORWARD;
  DevBuf diag;          // T [n] cached diagonal (src/gauss_seidel.rs:81)
  LevelSched fwd, bwd;  // lower / upper pattern
  DevBuf tmp;           // T [n]: forward result for the symmetric variant
  DevBuf barrier;       // grid barrier words
  int64_t bad_row = -1; // first row with a missing / tiny diagonal (src/gauss_seidel.rs:72-78)
};

template <typename T>
DiagOp<T>* diag_from_host(Ctx* ctx, int diag_dtype, const void* diag, int64_t n);
template <typename T>
DiagOp<T>* diag_from_csr(CsrMat<T>* A);
template <typename T>
GsOp<T>* gs_create(CsrMat<T>* A, int mode);
template <typename T>
void csr_diagonal(CsrMat<T>* A, T* d_diag);  // device out, 0 where absent

// out = M in for Diag / GS operators (device pointers).
template <typename T>
void diag_apply(DiagOp<T>* M, const T* in, T* out);
template <typename T>
void gs_apply(GsOp<T>* M, const T* in, T* out);
// One src/gauss_seidel.rs:111-125 sweep of the stationary solver: x_new from (x_old, rhs).
// Rows < i are read from x_new (already updated), rows > i from x_old.
template <typename T>
void gs_solver_sweep(GsOp<T>* M, const T* rhs, const T* x_old, T* x_new);

// Generic operator application used by the solvers: CSR -> SpMV, Diag, GS.
template <typename T>
void op_apply(spb_op* op, const T* in, T* out);

}  // namespace spb
