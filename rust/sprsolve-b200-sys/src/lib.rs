//! Rust binding of `libsprsolve_b200.so` with the interface of cxzheng/sprsolve.
//!
//! `GpuCsrMat<T>` is a drop-in for `MklMat<T>` / `CsMatI<T, i32>` wherever a `MatVecMul<T>` is
//! expected; `GpuBiCGStab` / `GpuMinRes` / `GpuCSMinRes` / `GpuGaussSeidel` mirror the reference's
//! solver types but run the whole loop on the device (one H2D of rhs/x, one D2H of x).
#![allow(non_snake_case, non_camel_case_types, clippy::many_single_char_names)]

use cauchy::Scalar;
use num_complex::{Complex32, Complex64};
use sprs::CsMatI;
use sprsolve::error::{SolveResult, SolverError};
use sprsolve::MatVecMul;
use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::{c_char, c_double, c_int, c_void};

// ------------------------------------------------------------------ raw ABI (include/sprsolve_b200.h)
#[repr(C)]
pub struct spb_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct spb_op {
    _p: [u8; 0],
}
#[repr(C)]
pub struct spb_solver {
    _p: [u8; 0],
}

pub const SPB_OK: c_int = 0;
pub const SPB_INCOMPATIBLE_FORMAT: c_int = 1;
pub const SPB_ZERO_DIAGONAL: c_int = 2;
pub const SPB_INSUFFICIENT_ITER: c_int = 3;
pub const SPB_BREAKDOWN: c_int = 4;
pub const SPB_INVALID_PRECOND: c_int = 5;
pub const SPB_DIM_MISMATCH: c_int = 6;
pub const SPB_UNIMPLEMENTED: c_int = 7;
pub const SPB_F64: c_int = 0;
pub const SPB_C128: c_int = 1;
pub const SPB_F32: c_int = 2;
pub const SPB_C64: c_int = 3;
pub const SPB_GS_FORWARD: c_int = 0;
pub const SPB_GS_SYMMETRIC: c_int = 1;

extern "C" {
    pub fn spb_last_error() -> *const c_char;
    pub fn spb_init(device: c_int, ctx: *mut *mut spb_ctx) -> c_int;
    pub fn spb_finalize(ctx: *mut spb_ctx) -> c_int;
    pub fn spb_csr_create(
        ctx: *mut spb_ctx, dtype: c_int, nrows: i64, ncols: i64, row_begin: i64, row_end: i64,
        indptr: *const c_void, indptr_bits: c_int, indices: *const i32, values: *const c_void,
        out: *mut *mut spb_op,
    ) -> c_int;
    pub fn spb_csc_create(
        ctx: *mut spb_ctx, dtype: c_int, nrows: i64, ncols: i64, indptr: *const c_void, indptr_bits: c_int,
        row_indices: *const i32, values: *const c_void, out: *mut *mut spb_op,
    ) -> c_int;
    pub fn spb_csr_create_from_triplets(
        ctx: *mut spb_ctx, dtype: c_int, nrows: i64, ncols: i64, nnz: i64, rows: *const i32, cols: *const i32,
        values: *const c_void, out: *mut *mut spb_op,
    ) -> c_int;
    pub fn spb_csr_read_matrix_market(ctx: *mut spb_ctx, dtype: c_int, path: *const c_char, out: *mut *mut spb_op) -> c_int;
    pub fn spb_csr_plan_info(mat: *mut spb_op, info: *mut i64) -> c_int; // [8]
    pub fn spb_gs_schedule_info(gs: *mut spb_op, info: *mut i64, stats: *mut i64, stats_cap: i64) -> c_int; // info [16]
    pub fn spb_csr_mv_hint(mat: *mut spb_op, ncalls: c_int) -> c_int;
    pub fn spb_csr_mv_and_dotmv_hint(mat: *mut spb_op, ncalls: c_int) -> c_int;
    pub fn spb_op_destroy(op: *mut spb_op) -> c_int;
    pub fn spb_op_mul_vec(op: *mut spb_op, v_in: *const c_void, n_in: i64, v_out: *mut c_void, n_out: i64) -> c_int;
    pub fn spb_op_mul_vec_dot(
        op: *mut spb_op, v_in: *const c_void, n_in: i64, v_out: *mut c_void, n_out: i64, out: *mut c_double,
    ) -> c_int;
    pub fn spb_diag_precond_create(
        ctx: *mut spb_ctx, dtype: c_int, diag_dtype: c_int, diag: *const c_void, n: i64, out: *mut *mut spb_op,
    ) -> c_int;
    pub fn spb_gs_precond_create(mat: *mut spb_op, mode: c_int, out: *mut *mut spb_op) -> c_int;
    pub fn spb_gs_precond_create_relaxed(mat: *mut spb_op, mode: c_int, omega: c_double, out: *mut *mut spb_op) -> c_int;
    pub fn spb_bicgstab_create(A: *mut spb_op, size: i64, out: *mut *mut spb_solver) -> c_int;
    pub fn spb_minres_create(A: *mut spb_op, size: i64, out: *mut *mut spb_solver) -> c_int;
    pub fn spb_csminres_create(A: *mut spb_op, size: i64, out: *mut *mut spb_solver) -> c_int;
    pub fn spb_gauss_seidel_create(A: *mut spb_op, out: *mut *mut spb_solver) -> c_int;
    pub fn spb_gauss_seidel_create_relaxed(A: *mut spb_op, omega: c_double, out: *mut *mut spb_solver) -> c_int;
    pub fn spb_solver_solve(
        s: *mut spb_solver, precond: *mut spb_op, rhs: *const c_void, n_rhs: i64, x: *mut c_void, n_x: i64,
        max_iter: i64, tol: c_double, iters: *mut i64, resid: *mut c_double, hist: *mut c_double, hist_cap: i64,
        hist_len: *mut i64,
    ) -> c_int;
    pub fn spb_solver_destroy(s: *mut spb_solver) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(spb_last_error()).to_string_lossy().into_owned() }
}

/// The instantiated scalar types: `f64` and `Complex64` (SURVEY.md section 8f lists f32/c32 as next).
/// The four `cauchy::Scalar` types the reference is generic over (it dispatches s / d / c / z,
/// src/mkl_mat.rs:68-71).  Scalars cross the C ABI as doubles; for the f32 types they hold floats.
pub trait GpuScalar: Scalar {
    const DTYPE: c_int;
    /// dtype code of `Self::Real` (a real diagonal for a complex system, `DiagPrecond<T, T::Real>`)
    const REAL_DTYPE: c_int;
    fn from_pair(p: [f64; 2]) -> Self;
    fn real_to_f64(r: Self::Real) -> f64;
    fn real_from_f64(v: f64) -> Self::Real;
}
impl GpuScalar for f64 {
    const DTYPE: c_int = SPB_F64;
    const REAL_DTYPE: c_int = SPB_F64;
    fn from_pair(p: [f64; 2]) -> Self {
        p[0]
    }
    fn real_to_f64(r: f64) -> f64 {
        r
    }
    fn real_from_f64(v: f64) -> f64 {
        v
    }
}
impl GpuScalar for Complex64 {
    const DTYPE: c_int = SPB_C128;
    const REAL_DTYPE: c_int = SPB_F64;
    fn from_pair(p: [f64; 2]) -> Self {
        Complex64::new(p[0], p[1])
    }
    fn real_to_f64(r: f64) -> f64 {
        r
    }
    fn real_from_f64(v: f64) -> f64 {
        v
    }
}
impl GpuScalar for f32 {
    const DTYPE: c_int = SPB_F32;
    const REAL_DTYPE: c_int = SPB_F32;
    fn from_pair(p: [f64; 2]) -> Self {
        p[0] as f32
    }
    fn real_to_f64(r: f32) -> f64 {
        r as f64
    }
    fn real_from_f64(v: f64) -> f32 {
        v as f32 // exact: the library hands back values already rounded to float
    }
}
impl GpuScalar for Complex32 {
    const DTYPE: c_int = SPB_C64;
    const REAL_DTYPE: c_int = SPB_F32;
    fn from_pair(p: [f64; 2]) -> Self {
        Complex32::new(p[0] as f32, p[1] as f32)
    }
    fn real_to_f64(r: f32) -> f64 {
        r as f64
    }
    fn real_from_f64(v: f64) -> f32 {
        v as f32
    }
}

// ------------------------------------------------------------------ context
/// One GPU + one CUDA stream.  Not `Send`/`Sync` (like `MklMat`, which holds a raw handle).
pub struct GpuContext {
    h: *mut spb_ctx,
}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self, u32> {
        let mut h = std::ptr::null_mut();
        let st = unsafe { spb_init(device as c_int, &mut h) };
        if st != SPB_OK {
            return Err(st as u32);
        }
        Ok(GpuContext { h })
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { spb_finalize(self.h) };
    }
}

// ------------------------------------------------------------------ GpuCsrMat: the MklMat drop-in
pub struct GpuCsrMat<'c, T: GpuScalar> {
    h: *mut spb_op,
    size: usize,
    _ctx: PhantomData<&'c GpuContext>,
    _t: PhantomData<T>,
}

impl<'c, T: GpuScalar> GpuCsrMat<'c, T> {
    /// `MklMat::new` (src/mkl_mat.rs:32-74): consumes a square CSR matrix with i32 indices,
    /// copies it to the device and analyses it.  `Err(status)` like the reference's `Err(u32)`.
    pub fn new(ctx: &'c GpuContext, m: CsMatI<T, i32>) -> Result<Self, u32> {
        assert!(m.is_csr());
        let (nrow, ncol) = (m.rows(), m.cols());
        assert_eq!(ncol, nrow);
        let (indptr, indices, data) = m.into_raw_storage();
        let mut h = std::ptr::null_mut();
        let st = unsafe {
            spb_csr_create(
                ctx.h, T::DTYPE, nrow as i64, ncol as i64, 0, nrow as i64,
                indptr.as_ptr() as *const c_void, 32, indices.as_ptr(), data.as_ptr() as *const c_void, &mut h,
            )
        };
        if st != SPB_OK {
            return Err(st as u32);
        }
        Ok(GpuCsrMat { h, size: nrow, _ctx: PhantomData, _t: PhantomData })
    }
    /// A CSC matrix as the operator: the CSC branch of `CsMatViewI::mul_vec` (src/mat.rs:130-142).
    /// Transposed on the device with a stable sort, so `mul_vec` keeps the accumulation order of
    /// the reference's column-by-column loop.
    pub fn from_csc(ctx: &'c GpuContext, m: CsMatI<T, i32>) -> Result<Self, u32> {
        assert!(m.is_csc());
        let (nrow, ncol) = (m.rows(), m.cols());
        let (indptr, indices, data) = m.into_raw_storage();
        let mut h = std::ptr::null_mut();
        let st = unsafe {
            spb_csc_create(ctx.h, T::DTYPE, nrow as i64, ncol as i64, indptr.as_ptr() as *const c_void, 32,
                           indices.as_ptr(), data.as_ptr() as *const c_void, &mut h)
        };
        if st != SPB_OK {
            return Err(st as u32);
        }
        Ok(GpuCsrMat { h, size: nrow, _ctx: PhantomData, _t: PhantomData })
    }
    /// `sprs::TriMat::to_csr` on the device (the reference's fixtures: tests/test_minres.rs:65-119):
    /// sorted by (row, column), duplicates summed in input order.
    pub fn from_triplets(ctx: &'c GpuContext, n: usize, rows: &[i32], cols: &[i32], data: &[T]) -> Result<Self, u32> {
        assert!(rows.len() == cols.len() && cols.len() == data.len());
        let mut h = std::ptr::null_mut();
        let st = unsafe {
            spb_csr_create_from_triplets(ctx.h, T::DTYPE, n as i64, n as i64, rows.len() as i64, rows.as_ptr(),
                                         cols.as_ptr(), data.as_ptr() as *const c_void, &mut h)
        };
        if st != SPB_OK {
            return Err(st as u32);
        }
        Ok(GpuCsrMat { h, size: n, _ctx: PhantomData, _t: PhantomData })
    }
    /// `MklMat::size` (src/mkl_mat.rs:26-28)
    #[inline(always)]
    pub fn size(&self) -> usize {
        self.size
    }
    /// `MklMat::mv_hint` (src/mkl_mat.rs:124-148)
    pub fn mv_hint(&self, ncalls: i32) -> Result<(), u32> {
        match unsafe { spb_csr_mv_hint(self.h, ncalls) } {
            SPB_OK => Ok(()),
            st => Err(st as u32),
        }
    }
    /// `MklMat::mv_and_dotmv_hint` (src/mkl_mat.rs:81-118)
    pub fn mv_and_dotmv_hint(&self, ncalls: i32) -> Result<(), u32> {
        match unsafe { spb_csr_mv_and_dotmv_hint(self.h, ncalls) } {
            SPB_OK => Ok(()),
            st => Err(st as u32),
        }
    }
}
impl<'c, T: GpuScalar> Drop for GpuCsrMat<'c, T> {
    fn drop(&mut self) {
        unsafe { spb_op_destroy(self.h) };
    }
}

macro_rules! impl_matvecmul {
    ($ty:ident) => {
        impl<'c, T: GpuScalar> MatVecMul<T> for $ty<'c, T> {
            fn mul_vec(&self, v_in: &[T], v_out: &mut [T]) {
                let st = unsafe {
                    spb_op_mul_vec(self.h, v_in.as_ptr() as *const c_void, v_in.len() as i64,
                                   v_out.as_mut_ptr() as *mut c_void, v_out.len() as i64)
                };
                match st {
                    SPB_OK => {}
                    SPB_DIM_MISMATCH => panic!("Dimension mismatch"), // src/mat.rs:50-52
                    _ => panic!("sprsolve-b200: {}", last_error()),   // src/mkl_mat.rs:188-193
                }
            }
            fn mul_vec_dot(&self, v_in: &[T], v_out: &mut [T]) -> T {
                let mut out = [0.0f64; 2];
                let st = unsafe {
                    spb_op_mul_vec_dot(self.h, v_in.as_ptr() as *const c_void, v_in.len() as i64,
                                       v_out.as_mut_ptr() as *mut c_void, v_out.len() as i64, out.as_mut_ptr())
                };
                match st {
                    SPB_OK => T::from_pair(out),
                    SPB_DIM_MISMATCH => panic!("Dimension mismatch"),
                    SPB_UNIMPLEMENTED => unimplemented!(), // src/precond.rs:55-62
                    _ => panic!("sprsolve-b200: {}", last_error()),
                }
            }
            // The ABI always checks sizes (a wrong size would be UB in the reference too).
            unsafe fn mul_vec_unchecked(&self, v_in: &[T], v_out: &mut [T]) {
                self.mul_vec(v_in, v_out)
            }
            unsafe fn mul_vec_dot_unchecked(&self, v_in: &[T], v_out: &mut [T]) -> T {
                self.mul_vec_dot(v_in, v_out)
            }
        }
    };
}
impl_matvecmul!(GpuCsrMat);

// ------------------------------------------------------------------ preconditioners
/// `DiagPrecond<T, V>` (src/precond.rs:6-63); `V` = `f64` or `T`.
pub struct GpuDiagPrecond<'c, T: GpuScalar> {
    h: *mut spb_op,
    _ctx: PhantomData<&'c GpuContext>,
    _t: PhantomData<T>,
}
impl<'c, T: GpuScalar> GpuDiagPrecond<'c, T> {
    pub fn new<V: GpuScalar>(ctx: &'c GpuContext, diag: &[V]) -> Self {
        let mut h = std::ptr::null_mut();
        let st = unsafe {
            spb_diag_precond_create(ctx.h, T::DTYPE, V::DTYPE, diag.as_ptr() as *const c_void,
                                    diag.len() as i64, &mut h)
        };
        assert_eq!(st, SPB_OK, "{}", last_error());
        GpuDiagPrecond { h, _ctx: PhantomData, _t: PhantomData }
    }
}
impl<'c, T: GpuScalar> Drop for GpuDiagPrecond<'c, T> {
    fn drop(&mut self) {
        unsafe { spb_op_destroy(self.h) };
    }
}
impl_matvecmul!(GpuDiagPrecond);

/// Level-scheduled Gauss-Seidel / symmetric Gauss-Seidel sweep as an operator (sweep body
/// src/gauss_seidel.rs:111-125).  No counterpart type exists in the reference.
pub struct GpuGsPrecond<'c, T: GpuScalar> {
    h: *mut spb_op,
    _ctx: PhantomData<&'c GpuContext>,
    _t: PhantomData<T>,
}
impl<'c, T: GpuScalar> GpuGsPrecond<'c, T> {
    pub fn new(A: &GpuCsrMat<'c, T>, symmetric: bool) -> SolveResult<Self> {
        Self::new_relaxed(A, symmetric, 1.0)
    }
    /// Relaxed sweep, 0 < omega < 2 (forward: SOR sweep from zero; symmetric: SSOR(omega)).
    pub fn new_relaxed(A: &GpuCsrMat<'c, T>, symmetric: bool, omega: f64) -> SolveResult<Self> {
        let mut h = std::ptr::null_mut();
        let mode = if symmetric { SPB_GS_SYMMETRIC } else { SPB_GS_FORWARD };
        match unsafe { spb_gs_precond_create_relaxed(A.h, mode, omega, &mut h) } {
            SPB_OK => Ok(GpuGsPrecond { h, _ctx: PhantomData, _t: PhantomData }),
            SPB_ZERO_DIAGONAL => {
                let msg = last_error();
                let row = msg.rsplit(' ').next().and_then(|s| s.parse().ok()).unwrap_or(0);
                Err(SolverError::ZeorDiagonalElem(row))
            }
            _ => Err(SolverError::IncompatibleMatrixFormat(last_error())),
        }
    }
}
impl<'c, T: GpuScalar> Drop for GpuGsPrecond<'c, T> {
    fn drop(&mut self) {
        unsafe { spb_op_destroy(self.h) };
    }
}
impl_matvecmul!(GpuGsPrecond);

/// Anything that lives on the device and can be handed to a device-resident solver.
pub trait GpuOp {
    fn raw(&self) -> *mut spb_op;
}
impl<'c, T: GpuScalar> GpuOp for GpuCsrMat<'c, T> {
    fn raw(&self) -> *mut spb_op {
        self.h
    }
}
impl<'c, T: GpuScalar> GpuOp for GpuDiagPrecond<'c, T> {
    fn raw(&self) -> *mut spb_op {
        self.h
    }
}
impl<'c, T: GpuScalar> GpuOp for GpuGsPrecond<'c, T> {
    fn raw(&self) -> *mut spb_op {
        self.h
    }
}

// ------------------------------------------------------------------ solvers
fn to_result<R>(st: c_int, iters: i64, resid: R) -> SolveResult<(usize, R)> {
    match st {
        SPB_OK => Ok((iters as usize, resid)),
        SPB_INCOMPATIBLE_FORMAT => Err(SolverError::IncompatibleMatrixFormat(last_error())),
        SPB_ZERO_DIAGONAL => Err(SolverError::ZeorDiagonalElem(iters as usize)),
        SPB_INSUFFICIENT_ITER => Err(SolverError::InsufficientIterNum(iters as usize)),
        SPB_BREAKDOWN => Err(SolverError::BreakDown(iters as usize)),
        SPB_INVALID_PRECOND => Err(SolverError::InvalidPreconditioner(last_error())),
        _ => panic!("sprsolve-b200: {}", last_error()),
    }
}

macro_rules! gpu_solver {
    ($name:ident, $create:ident, $doc:expr, precond = $has_pc:tt) => {
        #[doc = $doc]
        pub struct $name<'data, 'c, T: GpuScalar> {
            h: *mut spb_solver,
            _A: &'data GpuCsrMat<'c, T>,
        }
        impl<'data, 'c, T: GpuScalar> $name<'data, 'c, T> {
            /// Same signature as the reference's `new(A: &'data M, size: usize)`; allocates the
            /// device workspace once.
            pub fn new(A: &'data GpuCsrMat<'c, T>, size: usize) -> Self {
                let mut h = std::ptr::null_mut();
                let st = unsafe { $create(A.h, size as i64, &mut h) };
                assert_eq!(st, SPB_OK, "{}", last_error());
                $name { h, _A: A }
            }
            fn run(&mut self, pc: *mut spb_op, rhs: &[T], x: &mut [T], max_iter: usize, tol: T::Real)
                   -> SolveResult<(usize, T::Real)> {
                let (mut iters, mut resid) = (0i64, 0f64);
                let st = unsafe {
                    spb_solver_solve(self.h, pc, rhs.as_ptr() as *const c_void, rhs.len() as i64,
                                     x.as_mut_ptr() as *mut c_void, x.len() as i64, max_iter as i64, T::real_to_f64(tol),
                                     &mut iters, &mut resid, std::ptr::null_mut(), 0, std::ptr::null_mut())
                };
                to_result(st, iters, T::real_from_f64(resid))
            }
            /// Solves Ax = b, without preconditioner (the reference's `solve(&mut self, rhs, x, max_iter, tol: T::Real)
            /// -> SolveResult<(usize, T::Real)>`, src/bicg_stab.rs:35-41).
            pub fn solve(&mut self, rhs: &[T], x: &mut [T], max_iter: usize, tol: T::Real) -> SolveResult<(usize, T::Real)> {
                self.run(std::ptr::null_mut(), rhs, x, max_iter, tol)
            }
            gpu_solver!(@pc $has_pc);
        }
        impl<'data, 'c, T: GpuScalar> Drop for $name<'data, 'c, T> {
            fn drop(&mut self) {
                unsafe { spb_solver_destroy(self.h) };
            }
        }
    };
    (@pc yes) => {
        /// Same as the reference's `precond_solve<P: MatVecMul<T>>`, for device-resident `P`.
        pub fn precond_solve<P: GpuOp>(&mut self, precond: &P, rhs: &[T], x: &mut [T], max_iter: usize, tol: T::Real)
                                       -> SolveResult<(usize, T::Real)> {
            self.run(precond.raw(), rhs, x, max_iter, tol)
        }
    };
    (@pc no) => {};
}

gpu_solver!(GpuBiCGStab, spb_bicgstab_create, "`BiCGStab` (src/bicg_stab.rs:17-367)", precond = yes);
gpu_solver!(GpuMinRes, spb_minres_create, "`MinRes` (src/minres.rs:13-342)", precond = yes);
gpu_solver!(GpuCSMinRes, spb_csminres_create, "`CSMinRes` (src/cs_minres.rs:11-159)", precond = no);

/// `GaussSeidel` (src/gauss_seidel.rs:8-141): returns the ABSOLUTE residual.
pub struct GpuGaussSeidel<'data, 'c, T: GpuScalar> {
    h: *mut spb_solver,
    _A: &'data GpuCsrMat<'c, T>,
}
impl<'data, 'c, T: GpuScalar> GpuGaussSeidel<'data, 'c, T> {
    pub fn new(A: &'data GpuCsrMat<'c, T>) -> SolveResult<Self> {
        Self::new_relaxed(A, 1.0)
    }
    /// Successive over-relaxation, 0 < omega < 2 (omega = 1: the reference's solver).
    pub fn new_relaxed(A: &'data GpuCsrMat<'c, T>, omega: f64) -> SolveResult<Self> {
        let mut h = std::ptr::null_mut();
        match unsafe { spb_gauss_seidel_create_relaxed(A.h, omega, &mut h) } {
            SPB_OK => Ok(GpuGaussSeidel { h, _A: A }),
            _ => Err(SolverError::IncompatibleMatrixFormat(last_error())),
        }
    }
    pub fn solve(&mut self, rhs: &[T], x: &mut [T], max_iter: usize, eps: T::Real) -> SolveResult<(usize, T::Real)> {
        let (mut iters, mut resid) = (0i64, 0f64);
        let st = unsafe {
            spb_solver_solve(self.h, std::ptr::null_mut(), rhs.as_ptr() as *const c_void, rhs.len() as i64,
                             x.as_mut_ptr() as *mut c_void, x.len() as i64, max_iter as i64, T::real_to_f64(eps), &mut iters,
                             &mut resid, std::ptr::null_mut(), 0, std::ptr::null_mut())
        };
        to_result(st, iters, T::real_from_f64(resid))
    }
}
impl<'data, 'c, T: GpuScalar> Drop for GpuGaussSeidel<'data, 'c, T> {
    fn drop(&mut self) {
        unsafe { spb_solver_destroy(self.h) };
    }
}
