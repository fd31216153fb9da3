/* c_api_demo.c -- the drop-in boundary from plain C (C99): what a binding of the reference crate
 * would call, in the order the reference's own example does (src/main.rs:90-131: build the 2-D
 * Dirichlet Laplacian, solve with BiCGStab, print the iteration count).
 *
 *   gcc -std=c99 -Iinclude examples/c_api_demo.c -Lsprsolve_b200/lib -lsprsolve_b200 \
 *       -Wl,-rpath,$PWD/sprsolve_b200/lib -lm -o c_api_demo && ./c_api_demo [grid]
 *
 * Exit code 0 and a line "converged ..." on a B200; "no CUDA device" (exit 3) elsewhere: there is no
 * CPU fallback.  tests/test_abi.py compiles and links this file on every CPU run. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "sprsolve_b200.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int st_ = (call);                                                                 \
    if (st_ != SPB_OK) {                                                              \
      fprintf(stderr, "%s -> status %d: %s\n", #call, st_, spb_last_error());         \
      return st_ == SPB_NO_DEVICE ? 3 : 1;                                            \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  const int64_t g = argc > 1 ? atoll(argv[1]) : 96;
  const int64_t n = g * g;
  spb_ctx* ctx = NULL;
  int st = spb_init(0, &ctx);
  if (st == SPB_NO_DEVICE) {
    printf("no CUDA device: %s\n", spb_last_error());
    return 3;
  }
  CHECK(st);

  /* A: the reference's generator (src/main.rs:53-88), built on the device */
  spb_op* A = NULL;
  CHECK(spb_csr_create_stencil(ctx, SPB_STENCIL_DIRICHLET2D, SPB_F64, g, g, 1, NULL, 0, &A));
  int64_t nnz = 0, info[8];
  CHECK(spb_csr_nnz(A, &nnz));
  CHECK(spb_csr_plan_info(A, info));

  /* rhs: i + j on the border, 0 inside (src/main.rs:9-11, 90-103); x0 = 0 */
  double* rhs = (double*)calloc((size_t)n, sizeof(double));
  double* x = (double*)calloc((size_t)n, sizeof(double));
  double* y = (double*)calloc((size_t)n, sizeof(double));
  if (!rhs || !x || !y) return 2;
  for (int64_t i = 0; i < g; ++i)
    for (int64_t j = 0; j < g; ++j)
      if (i == 0 || j == 0 || i == g - 1 || j == g - 1) rhs[i * g + j] = (double)(i + j);

  /* MatVecMul::mul_vec on host slices, then DiagPrecond + BiCGStab::precond_solve */
  CHECK(spb_op_mul_vec(A, rhs, n, y, n));
  spb_op* M = NULL;
  CHECK(spb_diag_precond_from_csr(A, &M));
  spb_solver* S = NULL;
  CHECK(spb_bicgstab_create(A, n, &S));
  int64_t iters = 0, hist_len = 0;
  double resid = 0.0, hist[64];
  st = spb_solver_solve(S, M, rhs, n, x, n, 5000, 1e-8, &iters, &resid, hist, 64, &hist_len);
  if (st != SPB_OK) {
    fprintf(stderr, "solve -> status %d (%s), iters %lld\n", st, spb_last_error(), (long long)iters);
    return 1;
  }
  /* the harmonic function i + j solves the discrete problem exactly */
  double err = 0.0;
  for (int64_t i = 0; i < g; ++i)
    for (int64_t j = 0; j < g; ++j) err = fmax(err, fabs(x[i * g + j] - (double)(i + j)));
  printf("converged: n=%lld nnz=%lld iterations=%lld rel_residual=%.3e max_err=%.3e dictionary=%lld launches=%lld\n", (long long)n,
         (long long)nnz, (long long)iters, resid, err, (long long)info[0], (long long)spb_launch_count(ctx));

  CHECK(spb_solver_destroy(S));
  CHECK(spb_op_destroy(M));
  CHECK(spb_op_destroy(A));
  CHECK(spb_finalize(ctx));
  free(rhs);
  free(x);
  free(y);
  return err < 1e-4 ? 0 : 1;
}
