// ops.cuh -- preconditioner operators implementing MatVecMul<T>: DiagPrecond (src/precond.rs)
// and the level-scheduled Gauss-Seidel sweep operator (sweep body: src/gauss_seidel.rs:111-125).
#pragma once
#include "csr.cuh"

namespace spb {

template <typename T>
struct DiagOp : spb_op {
  DevBuf dinv;        // T or T::Real (real_diag) [n]: 1/diag, src/precond.rs:20-24
  bool real_diag = false;
};

// Level schedule of one triangular dependency pattern.
struct LevelSched {
  int64_t nlevels = 0;
  DevBuf level_ptr;   // int32 [nlevels+1]
  DevBuf rows;        // int32 [n]: rows sorted by level (ascending row id inside a level)
  std::vector<int> level_ptr_host;
};

// Block-wavefront schedule of one sweep direction (gs_wave.cu).  The rows are cut into blocks of
// `block_rows` consecutive rows, one CTA each; inside a block the rows are ordered by LOCAL level
// (dependencies on rows of the same block only) and packed, together with everything static a row
// needs, into a byte stream of chunks that the CTA's producer lane pulls through shared memory
// with bulk async copies.  Dependencies on rows of other blocks are resolved by polling the
// output vector (sentinel pre-fill), dependencies inside the block through shared memory.
struct WaveChunk {
  long long soff;     // byte offset of the chunk's static part in `stat`
  long long rhs_off;  // element offset of the chunk's slice of the permuted rhs
  long long aux_off;  // element offset of the chunk's slice of the per-apply other-triangle data
  int sbytes;         // static bytes (multiple of 16)
  int nrows;
  int aux_cnt;        // backward: nrows (pre-folded lower part); forward: Wo * nrows products
  int pad;
};

struct WaveSched {
  bool ok = false;
  bool backward = false;
  int block_rows = 0;
  int nblocks = 0;
  int64_t nchunks = 0;
  int64_t rhs_slots = 0, aux_slots = 0;
  int stage_static = 0, stage_rows = 0, stage_other = 0;  // per-stage capacities
  int stages = 0;
  size_t smem_bytes = 0;
  int64_t local_levels_max = 0;
  int64_t global_levels = 0;  // length of the longest dependency chain (levels of the whole triangle)
  DevBuf stat;       // packed static chunks
  DevBuf chunks;     // WaveChunk [nchunks]
  DevBuf blk_chunk;  // int32 [nblocks+1], blocks in PROCESSING order
  DevBuf rowmap;     // int32 [rhs_slots]: row id of the slot, -1 = padding
  DevBuf aux_base;   // int64 [rhs_slots] (forward only): first aux element of the row's products
  DevBuf aux_dims;   // int32 [2 * rhs_slots] (forward only): stride (rows of the chunk), Wo
  DevBuf rhsp;       // T [rhs_slots]  (per apply)
  DevBuf aux;        // T [aux_slots]  (per apply)
  DevBuf mailbox;    // T [mailbox_slots]: cross-block values, one contiguous slot range per consumer chunk (per apply)
  int64_t mailbox_slots = 0;
  DevBuf ticket;     // int32 [4]: block ticket, timeout flag
  int cluster = -1;  // CTAs per thread-block cluster of the sweep launch (0: none; -1: decided by the first sweep)
};

template <typename T>
struct GsOp : spb_op {
  CsrMat<T>* A = nullptr;
  int mode = SPB_GS_FORWARD;
  double omega = 1.0;   // relaxation factor: != 1 selects the relaxed sweep (SOR / SSOR(omega)), level-scheduled kernel
  DevBuf diag;          // T [n] cached diagonal (src/gauss_seidel.rs:81)
  LevelSched fwd, bwd;  // lower / upper pattern, global levels: the fallback sweep (built on first use)
  bool levels_ready = false;
  WaveSched wfwd, wbwd; // block-wavefront schedules (the fast path)
  DevBuf tmp;           // T [n]: forward result for the symmetric variant
  DevBuf sig;           // T [n]: the forward sweep's fold over the lower entries (the backward sweep starts from it)
  DevBuf barrier;       // grid barrier words
  DevBuf wave_stats;    // int64 [4 * nblocks] per-block clocks of the last wavefront sweep (SPB_GS_STATS=1)
  int64_t bad_row = -1; // first row with a missing / tiny diagonal (src/gauss_seidel.rs:72-78)
};

template <typename T>
DiagOp<T>* diag_from_host(Ctx* ctx, int diag_dtype, const void* diag, int64_t n);
template <typename T>
DiagOp<T>* diag_from_csr(CsrMat<T>* A);
template <typename T>
GsOp<T>* gs_create(CsrMat<T>* A, int mode, double omega = 1.0);
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

// gs_wave.cu: block-wavefront sweep.  wave_build analyses one direction (host CSR copy in);
// wave_sweep runs out = sweep(rhs) where the produced side reads `out` itself and the other
// triangle reads `other` (null: skipped, i.e. a sweep from zero).  out may alias rhs / other.
template <typename T>
void wave_build(CsrMat<T>* A, const std::vector<int64_t>& ip, const std::vector<int>& cols,
                const std::vector<T>& vals, bool backward, WaveSched& ws);
template <typename T>
void wave_sweep(GsOp<T>* M, WaveSched& ws, const T* rhs, const T* other, T* out, T* sig = nullptr);

// Generic operator application used by the solvers: CSR -> SpMV, Diag, GS.
template <typename T>
void op_apply(spb_op* op, const T* in, T* out);

}  // namespace spb
