//! Raw bindings to `include/nalgebra_b200.h` (host-pointer entry points + the device twins used by benchmarks).
//! Status codes: 0 ok, 1 not positive definite, 2 singular; negative = error (`na_last_error`).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_long, c_void};

pub const NA_OK: c_int = 0;
pub const NA_NOT_PD: c_int = 1;
pub const NA_SINGULAR: c_int = 2;
pub const NA_EINVAL: c_int = -1;
pub const NA_ECUDA: c_int = -2;
pub const NA_ENOMEM: c_int = -3;

extern "C" {
    pub fn na_init(device: c_int) -> c_int;
    pub fn na_shutdown() -> c_int;
    pub fn na_last_error() -> *const c_char;
    pub fn na_version() -> *const c_char;
    pub fn na_host_alloc_pinned(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn na_host_free_pinned(ptr: *mut c_void) -> c_int;
    pub fn na_set_tuning(key: *const c_char, value: c_long) -> c_int;

    // seam 1: matrixmultiply::dgemm / sgemm (src/base/blas_uninit.rs:298-313, 276-291)
    pub fn na_dgemm(m: usize, k: usize, n: usize, alpha: f64, a: *const f64, rsa: isize, csa: isize,
                    b: *const f64, rsb: isize, csb: isize, beta: f64, c: *mut f64, rsc: isize, csc: isize) -> c_int;
    pub fn na_sgemm(m: usize, k: usize, n: usize, alpha: f32, a: *const f32, rsa: isize, csa: isize,
                    b: *const f32, rsb: isize, csb: isize, beta: f32, c: *mut f32, rsc: isize, csc: isize) -> c_int;
    pub fn na_dgemv(m: usize, n: usize, alpha: f64, a: *const f64, rsa: isize, csa: isize,
                    x: *const f64, incx: isize, beta: f64, y: *mut f64, incy: isize) -> c_int;
    pub fn na_dsyrk_lower(n: usize, k: usize, alpha: f64, a: *const f64, rsa: isize, csa: isize,
                          beta: f64, c: *mut f64, ldc: usize) -> c_int;

    // seam 2: core-nalgebra layouts
    pub fn na_cholesky_f64(n: usize, a: *mut f64, lda: usize, use_sub: c_int, sub: f64, fail_col: *mut usize) -> c_int;
    pub fn na_cholesky_solve_f64(n: usize, l: *const f64, lda: usize, b: *mut f64, ldb: usize, nrhs: usize) -> c_int;
    pub fn na_lu_f64(m: usize, n: usize, a: *mut f64, lda: usize, swaps: *mut usize, nswaps: *mut usize) -> c_int;
    pub fn na_lu_solve_f64(n: usize, lu: *const f64, lda: usize, swaps: *const usize, nswaps: usize,
                           b: *mut f64, ldb: usize, nrhs: usize) -> c_int;
    pub fn na_qr_f64(m: usize, n: usize, a: *mut f64, lda: usize, diag: *mut f64) -> c_int;
    pub fn na_qr_q_f64(m: usize, n: usize, qr: *const f64, lda: usize, diag: *const f64, q: *mut f64, ldq: usize) -> c_int;
    pub fn na_qr_q_tr_mul_f64(m: usize, n: usize, qr: *const f64, lda: usize, diag: *const f64,
                              b: *mut f64, ldb: usize, nrhs: usize) -> c_int;
    pub fn na_qr_solve_f64(n: usize, qr: *const f64, lda: usize, diag: *const f64, b: *mut f64, ldb: usize, nrhs: usize) -> c_int;
    pub fn na_full_piv_lu_f64(m: usize, n: usize, a: *mut f64, lda: usize, p_swaps: *mut usize, np: *mut usize,
                              q_swaps: *mut usize, nq: *mut usize) -> c_int;
    pub fn na_col_piv_qr_f64(m: usize, n: usize, a: *mut f64, lda: usize, diag: *mut f64, p_swaps: *mut usize, np: *mut usize) -> c_int;
    pub fn na_hessenberg_f64(n: usize, a: *mut f64, lda: usize, subdiag: *mut f64) -> c_int;
    pub fn na_symmetric_tridiagonal_f64(n: usize, a: *mut f64, lda: usize, off_diagonal: *mut f64) -> c_int;
    pub fn na_bidiagonal_f64(m: usize, n: usize, a: *mut f64, lda: usize, diagonal: *mut f64, off_diagonal: *mut f64) -> c_int;
    pub fn na_tri_solve_f64(lower: c_int, trans: c_int, unit_diag: c_int, n: usize, t: *const f64, ldt: usize,
                            b: *mut f64, ldb: usize, nrhs: usize) -> c_int;

    // device twins used by benches (data stays in HBM)
    pub fn na_dgemm_dev(m: usize, k: usize, n: usize, alpha: f64, a: *const f64, rsa: isize, csa: isize,
                        b: *const f64, rsb: isize, csb: isize, beta: f64, c: *mut f64, rsc: isize, csc: isize,
                        stream: *mut c_void) -> c_int;
    pub fn na_cholesky_f64_dev(n: usize, a: *mut f64, lda: usize, use_sub: c_int, sub: f64, fail_col: *mut usize,
                               stream: *mut c_void) -> c_int;
    pub fn na_lu_f64_dev(m: usize, n: usize, a: *mut f64, lda: usize, swaps: *mut usize, nswaps: *mut usize,
                         stream: *mut c_void) -> c_int;
    pub fn na_qr_f64_dev(m: usize, n: usize, a: *mut f64, lda: usize, diag: *mut f64, stream: *mut c_void) -> c_int;
}
