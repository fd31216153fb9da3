/*
 * sprs_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal CPU restatement of the iterative-solve hot path of cxzheng/sprsolve
 * (Rust, v0.1.4) in its non-MKL configuration: sequential left folds, no FMA
 * contraction (build with -ffp-contract=off), num_complex 0.3 complex arithmetic.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libsprsolve_b200.so) never links or calls it.
 *
 * Pinning: checked against every known-answer test the reference holds for this path
 * (src/mat.rs:208-280, src/mkl_mat.rs:342-463, src/vecalg.rs:612-842 and doctests) and the
 * known-solution integration fixtures (tests/test_complex_solve.rs, test_complex_solve2.rs);
 * see tests/test_oracle_kats.py.  The reference itself cannot be compiled here (no
 * cargo/rustc, nightly-only crate, MKL + unvendored git deps), so there is no oracle/_ref.
 * CSMinRes is never executed by any reference test: its parity is UNPINNED by the reference.
 * The Gauss-Seidel *preconditioner* operators (forward / symmetric) have no reference
 * implementation; they are defined here from the gauss_seidel.rs:111-125 sweep body.
 *
 * All complex arrays are interleaved (re, im) doubles.  Column indices are int32,
 * row pointers int64.  Status codes mirror src/error.rs:7-22.
 */
#ifndef SPRS_ORACLE_H
#define SPRS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_OK = 0,
  ORC_INCOMPATIBLE_FORMAT = 1, /* SolverError::IncompatibleMatrixFormat */
  ORC_ZERO_DIAGONAL = 2,       /* SolverError::ZeorDiagonalElem(row) -> *iters = row */
  ORC_INSUFFICIENT_ITER = 3,   /* SolverError::InsufficientIterNum(max_iter) */
  ORC_BREAKDOWN = 4,           /* SolverError::BreakDown(its) -> *iters = its */
  ORC_INVALID_PRECOND = 5      /* SolverError::InvalidPreconditioner */
};

/* Preconditioner kinds understood by the oracle solvers. */
enum {
  ORC_PC_NONE = 0,
  ORC_PC_DIAG = 1,      /* DiagPrecond<T,T>: pc_data = diag (T), n entries (precond.rs:20-29) */
  ORC_PC_DIAG_REAL = 2, /* DiagPrecond<Complex,f64>: pc_data = real diag, n entries */
  ORC_PC_GS_FWD = 3,    /* z = one forward gauss_seidel.rs:111-125 sweep from z=0, rhs=r */
  ORC_PC_GS_SYM = 4,    /* forward sweep from 0 then the same body over rows n-1..0 */
  /* Relaxed variants (the reference has no relaxation, so the update is
   * DEFINED here: x_i <- (1 - w) x_i + w g_i with g_i the gauss_seidel.rs:123 value, computed as
   * mul_real(x_i, 1 - w) + mul_real(g_i, w)).  pc_data = &omega (one T::Real).  omega == 1 is GS_FWD / GS_SYM. */
  ORC_PC_SOR_FWD = 5,   /* relaxed forward sweep from z = 0 */
  ORC_PC_SSOR = 6       /* SSOR(w): relaxed forward sweep from 0, then the relaxed sweep over rows n-1..0 */
};

/* ---- SpMV: src/mat.rs:68-129 (CSR), :130-142 (CSC), :145-152 (mul_vec_dot) ---- */
void orc_spmv_d(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                const double* x, double* y);
void orc_spmv_z(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                const double* x, double* y);
/* rayon stand-in (mat.rs:85-107): OpenMP static row chunks >= 128 rows; same numerics. */
void orc_spmv_par_d(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                    const double* x, double* y);
void orc_spmv_par_z(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                    const double* x, double* y);
void orc_spmv_csc_d(int64_t nrows, int64_t ncols, const int64_t* indptr, const int32_t* idx,
                    const double* a, const double* x, double* y);
void orc_spmv_csc_z(int64_t nrows, int64_t ncols, const int64_t* indptr, const int32_t* idx,
                    const double* a, const double* x, double* y);
void orc_spmv_dot_d(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                    const double* x, double* y, double* out);
void orc_spmv_dot_z(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                    const double* x, double* y, double* out /* re,im */);

/* ---- vecalg fallbacks: src/vecalg.rs:556-605 ---- */
double orc_dot_d(int64_t n, const double* x, const double* y);
double orc_conj_dot_d(int64_t n, const double* x, const double* y);
double orc_norm2_d(int64_t n, const double* x);
void orc_dot_z(int64_t n, const double* x, const double* y, double* out);
void orc_conj_dot_z(int64_t n, const double* x, const double* y, double* out);
double orc_norm2_z(int64_t n, const double* x);
void orc_axpy_d(int64_t n, double a, const double* x, double* y);
void orc_axpby_d(int64_t n, double a, const double* x, double b, double* y);
void orc_scale_d(int64_t n, double a, double* x);
void orc_axpy_z(int64_t n, const double* a, const double* x, double* y);
void orc_axpby_z(int64_t n, const double* a, const double* x, const double* b, double* y);
void orc_scale_z(int64_t n, const double* a, double* x);
void orc_rscale_z(int64_t n, double a, double* x);
void orc_conj_z(int64_t n, const double* x, double* out);
/* f32 flavours used only to replay the reference's f32 KATs (vecalg.rs:764-790). */
void orc_axpy_s(int64_t n, float a, const float* x, float* y);
void orc_axpby_s(int64_t n, float a, const float* x, float b, float* y);
float orc_conj_dot_s(int64_t n, const float* x, const float* y);
/* Complex32 (c32) dot / conj_dot replay (vecalg.rs:654-720) */
void orc_dot_c(int64_t n, const float* x, const float* y, float* out);
void orc_conj_dot_c(int64_t n, const float* x, const float* y, float* out);

/* ---- preconditioner apply: src/precond.rs:20-29,48-52; GS sweep gauss_seidel.rs:111-125 ---- */
void orc_diag_apply_d(int64_t n, const double* diag, const double* in, double* out);
void orc_diag_apply_z(int64_t n, const double* diag, const double* in, double* out);
void orc_diag_apply_zd(int64_t n, const double* diag_real, const double* in, double* out);
int orc_gs_apply_d(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                   int symmetric, const double* in, double* out);
int orc_gs_apply_z(int64_t n, const int64_t* indptr, const int32_t* idx, const double* a,
                   int symmetric, const double* in, double* out);

/* ---- solvers.  workspace: caller-owned, persists across solves like the reference's
 *      Vec<T> (bicg_stab.rs:28 -> 7n, minres.rs:24 -> 8n, cs_minres.rs:22 -> 7n,
 *      gauss_seidel.rs:29 -> 2n), in units of T.
 *      size = the `size` given to ::new (dimension check bicg_stab.rs:44,49).
 *      hist (optional, may be NULL): per-iteration relative residual;
 *        BiCGStab: hist[0] = ||r0||/||b|| (bicg_stab.rs:251), hist[its] = r_norm/rhs_norm at
 *                  the top of iteration its (bicg_stab.rs:296);
 *        MINRES / CSMINRES: hist[its] = res_norm/rhs_norm after iteration its (minres.rs:164).
 *      *hist_len receives the number of entries the solver produced (may exceed hist_cap;
 *      only the first hist_cap are stored).  */
int orc_bicgstab_d(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, int pc_kind, const double* pc_data,
                   const double* rhs, double* x, int64_t max_iter, double tol, double* work,
                   int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                   int64_t* hist_len);
int orc_bicgstab_z(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, int pc_kind, const double* pc_data,
                   const double* rhs, double* x, int64_t max_iter, double tol, double* work,
                   int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                   int64_t* hist_len);
int orc_minres_d(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                 const int32_t* idx, const double* a, int pc_kind, const double* pc_data,
                 const double* rhs, double* x, int64_t max_iter, double tol, double* work,
                 int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                 int64_t* hist_len);
int orc_minres_z(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                 const int32_t* idx, const double* a, int pc_kind, const double* pc_data,
                 const double* rhs, double* x, int64_t max_iter, double tol, double* work,
                 int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                 int64_t* hist_len);
int orc_csminres_d(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, const double* rhs, double* x,
                   int64_t max_iter, double tol, double* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len);
int orc_csminres_z(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                   const int32_t* idx, const double* a, const double* rhs, double* x,
                   int64_t max_iter, double tol, double* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len);
/* GaussSeidel::solve, gauss_seidel.rs:33-140 (returns the ABSOLUTE residual).
 * hist[k] = absolute residual after sweep k (k = 0 is the unrolled first sweep). */
int orc_gauss_seidel_d(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
                       const int64_t* indptr, const int32_t* idx, const double* a,
                       const double* rhs, double* x, int64_t max_iter, double eps, double* work,
                       int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                       int64_t* hist_len);
int orc_gauss_seidel_z(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,
                       const int64_t* indptr, const int32_t* idx, const double* a,
                       const double* rhs, double* x, int64_t max_iter, double eps, double* work,
                       int64_t* iters, double* resid, double* hist, int64_t hist_cap,
                       int64_t* hist_len);

/* ---- synthetic matrix generators.  Two-pass: call with
 *      indptr/idx/a == NULL to obtain nnz, then with buffers.  Sorted columns per row.  ---- */
/* reference generator src/main.rs:53-88 (Dirichlet identity rows, interior [1,1,-4,1,1]);
 * rhs = i+j on the border (main.rs:90-103).  Note the reference indexes i*rows+j. */
int64_t orc_gen_dirichlet2d(int64_t rows, int64_t cols, int64_t* indptr, int32_t* idx, double* a,
                            double* rhs);
/* 3-D 7-point: diag = 6 - shift (complex: 6 - (shift_re + i shift_im)), off-diagonals -1,
 * x fastest, truncated at the boundary. */
int64_t orc_gen_lap3d7_d(int64_t nx, int64_t ny, int64_t nz, double shift, int64_t* indptr,
                         int32_t* idx, double* a);
int64_t orc_gen_lap3d7_z(int64_t nx, int64_t ny, int64_t nz, double shift_re, double shift_im,
                         int64_t* indptr, int32_t* idx, double* a);
/* 27-point convection-diffusion: (27 I - S27) + bx Dx + by Dy + bz Dz, upwind bidiagonals
 * (sub = -1, diag = +1): centre 26 + bx + by + bz, face neighbours at x-1,y-1,z-1 get
 * -1 - b*, the other neighbours -1.  Rows [row_begin,row_end) only (for partitioned tests);
 * indptr is relative to row_begin. */
int64_t orc_gen_convdiff27_d(int64_t nx, int64_t ny, int64_t nz, double bx, double by, double bz,
                             int64_t row_begin, int64_t row_end, int64_t* indptr, int32_t* idx,
                             double* a);

/* Relaxed Gauss-Seidel operator (symmetric: SSOR(omega)) and the stationary SOR solver (GaussSeidel::solve
 * with the relaxed update; same returns).  Suffixes d / z / s / c; complex arrays interleaved. */
#define ORC_SOR_DECL(sfx, F)                                                                                          \
  int orc_sor_apply_##sfx(int64_t n, const int64_t* indptr, const int32_t* idx, const F* a, int symmetric,           \
                          double omega, const F* in, F* out);                                                         \
  int orc_sor_solve_##sfx(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x,                       \
                          const int64_t* indptr, const int32_t* idx, const F* a, const F* rhs, F* x, int64_t max_iter, \
                          double eps, double omega, F* work, int64_t* iters, double* resid, double* hist,             \
                          int64_t hist_cap, int64_t* hist_len);
ORC_SOR_DECL(d, double)
ORC_SOR_DECL(z, double)
ORC_SOR_DECL(s, float)
ORC_SOR_DECL(c, float)

int orc_max_threads(void);
void orc_set_threads(int n);
/* 0 = serial everything (the bit-defining oracle = reference without `parallel`/`mkl`);
 * 1 = OpenMP row-parallel SpMV only (= the reference's rayon `parallel` feature; same numerics);
 * 2 = OpenMP SpMV + OpenMP vector ops (stand-in for the `mkl` iomp build; summation order differs).
 * Modes 1/2 exist for the timed CPU baseline only. */
void orc_set_mode(int mode);
int orc_get_mode(void);

/* ---- f32 / Complex32 (float arrays, interleaved re,im for _c; all real-valued solver quantities in
 * float, thresholds with f32::EPSILON).  Restated for SURVEY.md section 8f rank 2; pinned by the
 * reference's f32 / c32 vecalg KATs (src/vecalg.rs:647-677, doctests :122-132), otherwise by
 * agreement with the f64 restatement to float accuracy (tests/test_oracle_f32.py). */
void orc_spmv_s(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, const float* x, float* y);
void orc_spmv_c(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, const float* x, float* y);
float orc_spmv_dot_s(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, const float* x, float* y);
void orc_spmv_dot_c(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, const float* x, float* y, float* out);
float orc_norm2_s(int64_t n, const float* x);
float orc_norm2_c(int64_t n, const float* x);
float orc_dot_s(int64_t n, const float* x, const float* y);
void orc_scale_s(int64_t n, float a, float* x);
void orc_scale_c(int64_t n, const float* a, float* x);
void orc_rscale_s(int64_t n, float a, float* x);
void orc_rscale_c(int64_t n, float a, float* x);
void orc_axpy_c(int64_t n, const float* a, const float* x, float* y);
void orc_axpby_c(int64_t n, const float* a, const float* x, const float* b, float* y);
void orc_conj_c(int64_t n, const float* x, float* out);
void orc_diag_apply_s(int64_t n, const float* diag, const float* in, float* out);
void orc_diag_apply_c(int64_t n, const float* diag, const float* in, float* out);
void orc_diag_apply_cs(int64_t n, const float* diag, const float* in, float* out);
int orc_gs_apply_s(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, int symmetric, const float* in, float* out);
int orc_gs_apply_c(int64_t n, const int64_t* indptr, const int32_t* idx, const float* a, int symmetric, const float* in, float* out);
#define ORC_SOLVER_ARGS_F                                                                        \
  int64_t size, int64_t n_rhs, int64_t n_x, const int64_t *indptr, const int32_t *idx,          \
      const float *a, int pc_kind, const float *pc_data, const float *rhs, float *x,            \
      int64_t max_iter, double tol, float *work, int64_t *iters, double *resid, double *hist,   \
      int64_t hist_cap, int64_t *hist_len
int orc_bicgstab_s(ORC_SOLVER_ARGS_F);
int orc_bicgstab_c(ORC_SOLVER_ARGS_F);
int orc_minres_s(ORC_SOLVER_ARGS_F);
int orc_minres_c(ORC_SOLVER_ARGS_F);
int orc_csminres_s(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr, const int32_t* idx, const float* a,
                   const float* rhs, float* x, int64_t max_iter, double tol, float* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len);
int orc_csminres_c(int64_t size, int64_t n_rhs, int64_t n_x, const int64_t* indptr, const int32_t* idx, const float* a,
                   const float* rhs, float* x, int64_t max_iter, double tol, float* work, int64_t* iters, double* resid,
                   double* hist, int64_t hist_cap, int64_t* hist_len);
int orc_gauss_seidel_s(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                       const int32_t* idx, const float* a, const float* rhs, float* x, int64_t max_iter, double eps,
                       float* work, int64_t* iters, double* resid, double* hist, int64_t hist_cap, int64_t* hist_len);
int orc_gauss_seidel_c(int64_t nrows, int64_t ncols, int is_csr, int64_t n_rhs, int64_t n_x, const int64_t* indptr,
                       const int32_t* idx, const float* a, const float* rhs, float* x, int64_t max_iter, double eps,
                       float* work, int64_t* iters, double* resid, double* hist, int64_t hist_cap, int64_t* hist_len);

#ifdef __cplusplus
}
#endif
#endif
