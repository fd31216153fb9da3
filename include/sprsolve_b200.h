/*
 * sprsolve_b200.h -- C ABI of libsprsolve_b200.so: a B200 (sm_100a) drop-in for the iterative-solve
 * hot path of cxzheng/sprsolve (Rust).  Plain pointers and sizes only; no C++/torch types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * crate root).  The reference-side binding a maintainer would add (a Rust `extern "C"` block plus
 * `impl MatVecMul<T> for GpuCsrMat<T>`) is shown in INTEGRATION.md and shipped as source in rust/.
 *
 * Conventions
 *   - dtype: SPB_F64 = f64, SPB_C128 = num_complex::Complex64 (interleaved re,im doubles),
 *     SPB_F32 = f32, SPB_C64 = num_complex::Complex32 (interleaved re,im floats).  Scalars that cross
 *     the ABI (a, b, dot results, tol, resid) are always doubles; for the f32 types they hold floats.
 *   - Column indices are int32 (MklMat: `Vec<i32>`, src/mkl_mat.rs:17-18); row pointers int32 or int64.
 *   - Every function returns an spb_status.  Values 1..5 map 1:1 onto `SolverError`
 *     (src/error.rs:7-22).  Nothing panics or throws across the ABI; where the reference panics
 *     ("Dimension mismatch", src/mat.rs:50-52) the call returns SPB_DIM_MISMATCH, and where it hits
 *     `unimplemented!()` (src/precond.rs:55-62) it returns SPB_UNIMPLEMENTED.
 *   - There is NO CPU fallback: without a usable CUDA device spb_init fails with SPB_NO_DEVICE.
 *   - Handles are opaque, owned by the library, freed by the matching *_destroy.  One solve at a
 *     time per context (the reference's solvers take `&mut self`).
 *   - "host" entry points take caller-owned host slices exactly like the reference's `&[T]` /
 *     `&mut [T]` arguments (H2D / D2H inside the call); "_dev" entry points take device pointers.
 *   - Multi-GPU: one process (one spb_ctx) per GPU; the matrix is partitioned by contiguous row
 *     blocks, vectors passed to a distributed operator/solver are the LOCAL slices.
 */
#ifndef SPRSOLVE_B200_H
#define SPRSOLVE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct spb_ctx spb_ctx;       /* device + stream + (optional) NCCL communicator          */
typedef struct spb_op spb_op;         /* anything that implements MatVecMul<T> (src/mat.rs:12-37) */
typedef struct spb_solver spb_solver; /* BiCGStab / MinRes / CSMinRes / GaussSeidel + workspace   */

typedef enum {
  SPB_OK = 0,
  SPB_INCOMPATIBLE_FORMAT = 1, /* SolverError::IncompatibleMatrixFormat (src/error.rs:8-9)          */
  SPB_ZERO_DIAGONAL = 2,       /* SolverError::ZeorDiagonalElem(row): *iters receives the row      */
  SPB_INSUFFICIENT_ITER = 3,   /* SolverError::InsufficientIterNum(max_iter)                       */
  SPB_BREAKDOWN = 4,           /* SolverError::BreakDown(its): *iters receives its                 */
  SPB_INVALID_PRECOND = 5,     /* SolverError::InvalidPreconditioner                               */
  SPB_DIM_MISMATCH = 6,        /* panic!("Dimension mismatch") of mul_vec / mul_vec_dot            */
  SPB_UNIMPLEMENTED = 7,       /* unimplemented!() of DiagPrecond::mul_vec_dot                     */
  SPB_CUDA_ERROR = 100,
  SPB_NCCL_ERROR = 101,
  SPB_INVALID_ARG = 102,
  SPB_NO_DEVICE = 103
} spb_status;

/* cauchy::Scalar is implemented for f32, f64, Complex32, Complex64; MklMat dispatches s/d/c/z
 * (src/mkl_mat.rs:68-71,198-201,220).  The f32 types compute every T::Real quantity in float. */
typedef enum { SPB_F64 = 0, SPB_C128 = 1, SPB_F32 = 2, SPB_C64 = 3 } spb_dtype;

/* Synthetic on-device matrix generators.  params: see each kind. */
typedef enum {
  SPB_STENCIL_DIRICHLET2D = 0, /* src/main.rs:53-88; nx = rows, ny = cols (square); nz ignored      */
  SPB_STENCIL_LAP3D7 = 1,      /* 7-point, diag 6 - shift, off -1; params = {shift_re, shift_im}   */
  SPB_STENCIL_CONVDIFF27 = 2   /* 27-point convection-diffusion; params = {bx, by, bz}             */
} spb_stencil;

typedef enum { SPB_GS_FORWARD = 0, SPB_GS_SYMMETRIC = 1 } spb_gs_mode;

const char* spb_version(void);
/* Message of the last failure on the calling thread ("" if none). */
const char* spb_last_error(void);

/* ---- context -------------------------------------------------------------------------------- */
int spb_init(int device, spb_ctx** ctx);
int spb_finalize(spb_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (e.g. torch's current stream). */
int spb_set_stream(spb_ctx* ctx, void* cuda_stream);
int spb_synchronize(spb_ctx* ctx);
/* Number of kernels this library has launched on ctx so far (bench.py's gpu_launches). */
int64_t spb_launch_count(spb_ctx* ctx);
/* Per-kernel-family launch counts and accumulated CUDA-event time (only while profiling is on).
 * family: 0 = spmv, 1 = vector/fused, 2 = scalar, 3 = precond, 4 = halo pack. */
int spb_profile_enable(spb_ctx* ctx, int on);
int spb_profile_read(spb_ctx* ctx, int family, int64_t* launches, double* ms);
int spb_profile_reset(spb_ctx* ctx);

/* Multi-GPU (one process per GPU).  id is an opaque 128-byte NCCL unique id created on rank 0
 * with spb_comm_unique_id and distributed by the caller (torch.distributed / MPI / files). */
int spb_comm_unique_id(void* id128);
int spb_comm_init(spb_ctx* ctx, int world, int rank, const void* id128);
int spb_comm_info(spb_ctx* ctx, int* world, int* rank);
/* The row block [row_begin,row_end) spb_csr_create_stencil gives rank `rank` of `world`: contiguous
 * rows cut at plane boundaries (z-slabs).  Pure host arithmetic (usable without a GPU). */
int spb_stencil_partition(int kind, int64_t nx, int64_t ny, int64_t nz, int world, int rank,
                          int64_t* row_begin, int64_t* row_end);

/* ---- matrices: replaces MklMat::new (src/mkl_mat.rs:32-74) and the CsMatI operator ---------- */
/* Copies a host CSR matrix to the device and analyses it (tiling = the mkl_sparse_optimize
 * analogue).  indptr_bits is 32 or 64.  Must be square (src/mkl_mat.rs:37) else
 * SPB_INCOMPATIBLE_FORMAT.  With a communicator: [row_begin,row_end) is this rank's row block,
 * indptr is local (starts at 0), indices are GLOBAL columns, nrows is the global size. */
int spb_csr_create(spb_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t row_begin,
                   int64_t row_end, const void* indptr, int indptr_bits, const int32_t* indices,
                   const void* values, spb_op** out);
/* Generates the matrix on the device (this rank's row block when a communicator is set). */
int spb_csr_create_stencil(spb_ctx* ctx, int kind, int dtype, int64_t nx, int64_t ny, int64_t nz,
                           const double* params, int nparams, spb_op** out);

/* ---- other layouts on the callers' side of the operator (single GPU; all arrays are host) ------
 * CSC: the reference's CsMatViewI operator also accepts CSC storage and multiplies column by column,
 * v_out[row] += v_in[col] * value (src/mat.rs:130-142; KAT src/mat.rs:208-229).  The matrix is
 * transposed on the device with a stable sort, which keeps every output element's accumulation
 * order, so mul_vec is bit-identical to that loop.  indptr has ncols+1 entries. */
int spb_csc_create(spb_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, const void* indptr,
                   int indptr_bits, const int32_t* row_indices, const void* values, spb_op** out);
/* Triplets: sprs::TriMat::to_csr as used by the reference's fixtures (tests/test_minres.rs:65-119,
 * tests/test_complex_solve.rs:99-213): sorted by (row, column), duplicates summed in input order. */
int spb_csr_create_from_triplets(spb_ctx* ctx, int dtype, int64_t nrows, int64_t ncols, int64_t nnz,
                                 const int32_t* rows, const int32_t* cols, const void* values,
                                 spb_op** out);
/* Matrix Market coordinate file (real / integer / pattern / complex; general / symmetric /
 * skew-symmetric / hermitian), the format the reference's fixtures were exported from
 * (tests/test_complex_solve.rs:14).  Non-square or array-format files: SPB_INCOMPATIBLE_FORMAT. */
int spb_csr_read_matrix_market(spb_ctx* ctx, int dtype, const char* path, spb_op** out);
/* MklMat::mv_hint / mv_and_dotmv_hint (src/mkl_mat.rs:81-148), i.e. mkl_sparse_set_mv_hint +
 * mkl_sparse_optimize: time a handful of launch plans with real SpMV launches on this matrix and
 * keep the fastest.  Optional: spb_csr_create already picks a static plan.  Results never depend on
 * the plan (every row is folded in CSR order; the fused dot products are exactly rounded,
 * csrc/reduce.cuh), so it may be called at any time.
 * Collective when the matrix is partitioned (every rank must call it). */
int spb_csr_mv_hint(spb_op* mat, int ncalls);
int spb_csr_mv_and_dotmv_hint(spb_op* mat, int ncalls);
/* MklMat::size (src/mkl_mat.rs:26-28), plus local sizes for partitioned matrices. */
int spb_op_size(spb_op* op, int64_t* n_global, int64_t* n_local, int64_t* row_begin);
int spb_csr_nnz(spb_op* mat, int64_t* nnz_local);
/* What the analysis (the mkl_sparse_optimize analogue, src/mkl_mat.rs:104-121) decided for this matrix:
 * info = {column-offset dictionary on, distinct row patterns, dictionary row stride, consumer threads per CTA,
 * ring stages, non-zeros per tile, resident CTAs per SM, bytes one mul_vec streams from HBM by design}. */
int spb_csr_plan_info(spb_op* mat, int64_t info[8]);
/* Copy the local CSR arrays back (tests / generators parity).  indptr64 has n_local+1 entries. */
int spb_csr_download(spb_op* mat, int64_t* indptr64, int32_t* indices_global, void* values);
/* diag[i] = a_ii (0 when absent) of the local rows, host buffer of n_local T. */
int spb_csr_diagonal(spb_op* mat, void* diag_host);
int spb_op_destroy(spb_op* op);

/* ---- trait MatVecMul<T> (src/mat.rs:12-37) on host slices ------------------------------------ */
/* mul_vec: v_out = A v_in; checks dimensions like src/mat.rs:49-56 -> SPB_DIM_MISMATCH. */
int spb_op_mul_vec(spb_op* op, const void* v_in, int64_t n_in, void* v_out, int64_t n_out);
/* mul_vec_dot: v_out = A v_in, *out = conj(v_in) . v_out (src/mat.rs:59-64, mkl_mat.rs:242-319). */
int spb_op_mul_vec_dot(spb_op* op, const void* v_in, int64_t n_in, void* v_out, int64_t n_out,
                       double out[2]);
/* Same on device pointers (the "unchecked" variants: src/mat.rs:68, :145). */
int spb_op_mul_vec_dev(spb_op* op, const void* d_in, void* d_out);
int spb_op_mul_vec_dot_dev(spb_op* op, const void* d_in, void* d_out, double out[2]);

/* ---- preconditioners -------------------------------------------------------------------------- */
/* DiagPrecond::new(diag) (src/precond.rs:20-29): stores 1/diag.  diag_dtype may be SPB_F64 while
 * dtype is SPB_C128 (DiagPrecond<Complex64,f64>, tests/test_complex_solve.rs:44); likewise SPB_F32
 * for SPB_C64. */
int spb_diag_precond_create(spb_ctx* ctx, int dtype, int diag_dtype, const void* diag, int64_t n,
                            spb_op** out);
/* Same, taking diag(A) on the device (for generated matrices too large to stage on the host). */
int spb_diag_precond_from_csr(spb_op* mat, spb_op** out);
/* Level-scheduled Gauss-Seidel operator built from the src/gauss_seidel.rs:111-125 sweep body:
 * FORWARD: one sweep from z = 0 with rhs = v_in; SYMMETRIC: followed by the same body over rows
 * n-1..0.  Sequential update order preserved (thread-per-row accumulation in CSR order).
 * SPB_ZERO_DIAGONAL like src/gauss_seidel.rs:72-78.  Single-GPU only. */
int spb_gs_precond_create(spb_op* mat, int mode, spb_op** out);
/* Relaxed variant (the reference has no relaxation): every update of the
 * sweep becomes x_i <- (1 - omega) x_i + omega g_i, g_i the src/gauss_seidel.rs:123 value, 0 < omega < 2.
 * FORWARD: one SOR sweep from z = 0; SYMMETRIC: SSOR(omega) = omega (2 - omega) (D + omega U)^-1 D
 * (D + omega L)^-1, symmetric positive definite whenever A is symmetric with D > 0 (a valid MINRES
 * preconditioner, src/minres.rs:176).  omega == 1 is spb_gs_precond_create.  Sequential update order
 * preserved (level-scheduled kernel).  Single-GPU only. */
int spb_gs_precond_create_relaxed(spb_op* mat, int mode, double omega, spb_op** out);
int spb_gs_levels(spb_op* gs, int64_t* n_levels_fwd, int64_t* n_levels_bwd);
/* Block-wavefront schedule of the sweep (diagnostics): info = {fwd ok, rows per block, blocks, fwd chunks,
 * fwd local levels (max over blocks), ring stages, shared memory bytes, bwd ok, bwd chunks, bwd local levels,
 * poll-timeout flag, rhs slots, other-side slots, packed bytes, 0, 0}.  stats (optional, SPB_GS_STATS=1 at
 * create): per block of the last sweep {clocks, clocks waiting for the ring, shared-memory spins on
 * values of other blocks, clocks of thread 0 in the level barrier}. */
int spb_gs_schedule_info(spb_op* gs, int64_t info[16], int64_t* stats, int64_t stats_cap);

/* ---- vecalg (src/vecalg.rs:19-144) on host slices; out / a / b are (re,im) pairs -------------- */
int spb_vec_dot(spb_ctx* ctx, int dtype, int64_t n, const void* x, const void* y, double out[2]);
int spb_vec_conj_dot(spb_ctx* ctx, int dtype, int64_t n, const void* x, const void* y,
                     double out[2]);
int spb_vec_norm2(spb_ctx* ctx, int dtype, int64_t n, const void* x, double* out);
int spb_vec_scale(spb_ctx* ctx, int dtype, int64_t n, const double a[2], void* x);
int spb_vec_rscale(spb_ctx* ctx, int dtype, int64_t n, double a, void* x);
int spb_vec_conj(spb_ctx* ctx, int dtype, int64_t n, const void* x, void* out);
int spb_vec_axpy(spb_ctx* ctx, int dtype, int64_t n, const double a[2], const void* x, void* y);
int spb_vec_axpby(spb_ctx* ctx, int dtype, int64_t n, const double a[2], const void* x,
                  const double b[2], void* y);

/* ---- solvers ---------------------------------------------------------------------------------- */
/* BiCGStab::new(&A, size) (src/bicg_stab.rs:25-31), MinRes::new (src/minres.rs:21-27),
 * CSMinRes::new (src/cs_minres.rs:19-25): allocate the 7n / 8n / 7n device workspace once.
 * GaussSeidel::new (src/gauss_seidel.rs:15-31): SPB_INCOMPATIBLE_FORMAT unless square CSR. */
int spb_bicgstab_create(spb_op* A, int64_t size, spb_solver** out);
int spb_minres_create(spb_op* A, int64_t size, spb_solver** out);
int spb_csminres_create(spb_op* A, int64_t size, spb_solver** out);
int spb_gauss_seidel_create(spb_op* A, spb_solver** out);
/* The same stationary solver with successive over-relaxation (SOR), 0 < omega < 2; omega == 1 is
 * GaussSeidel::solve itself. */
int spb_gauss_seidel_create_relaxed(spb_op* A, double omega, spb_solver** out);
/* solve / precond_solve (src/bicg_stab.rs:35,204; src/minres.rs:31,178; src/cs_minres.rs:29;
 * src/gauss_seidel.rs:33).  precond == NULL selects `solve`.  x is in/out (initial guess).
 * On SPB_OK: *iters, *resid as the reference's Ok((iters, resid)).  hist (optional): see
 * oracle/sprs_oracle.h for the per-solver meaning; *hist_len = entries produced.
 * The whole loop runs on the device: one H2D of rhs/x, one D2H of x. */
int spb_solver_solve(spb_solver* s, spb_op* precond, const void* rhs, int64_t n_rhs, void* x,
                     int64_t n_x, int64_t max_iter, double tol, int64_t* iters, double* resid,
                     double* hist, int64_t hist_cap, int64_t* hist_len);
/* Same with rhs / x already resident in device memory (n = local size). */
int spb_solver_solve_dev(spb_solver* s, spb_op* precond, const void* d_rhs, void* d_x,
                         int64_t max_iter, double tol, int64_t* iters, double* resid, double* hist,
                         int64_t hist_cap, int64_t* hist_len);
/* Iterations between two host polls of the device-side status word (default 16). */
int spb_solver_set_poll_interval(spb_solver* s, int iters);
int spb_solver_destroy(spb_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* SPRSOLVE_B200_H */
