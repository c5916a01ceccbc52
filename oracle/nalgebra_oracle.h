/*
 * nalgebra_oracle.h -- CPU restatement of dimforge/nalgebra v0.35.0's dense hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a line-faithful, single-threaded C
 * restatement of the reference's algorithms (operation order included), used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the CHECKER and
 * the CPU baseline.  Nothing under nalgebra_b200/ (the product) may link, import or call it.
 *
 * The reference is Rust; there is no Rust toolchain in this image, so oracle/_ref (the reference
 * compiled from source) does not exist.  Pinning: the oracle is checked against every known-answer
 * test the reference's own test-suite holds for this path (tests/golden/nalgebra_kats.json, see
 * tests/test_oracle_golden.py).  One piece stays "parity unpinned": values at the matrixmultiply
 * boundary (third-party crate `matrixmultiply = "0.3"`, Cargo.toml:95, not vendored, no lockfile)
 * -- the reference has no known-answer test for f64 GEMM with every dim > 5.
 *
 * All matrices are column-major.  Strides are in elements.
 */
#ifndef NALGEBRA_ORACLE_H
#define NALGEBRA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- synthetic inputs (shared bit-for-bit with the CUDA generator and the numpy one) ---- */
double na_oracle_rand01(uint64_t seed, uint64_t idx);
void na_oracle_fill_uniform(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed);

/* ---- GEMM: src/base/blas_uninit.rs:187-333 (gemm_uninit) ---- */
/* C <- alpha*A*B + beta*C with nalgebra's dispatch: every dim > 5 -> matrixmultiply-style blocked
 * kernel (restated from the crate's published BLIS-like algorithm), else the gemv/axcpy fallback
 * in the reference's exact arithmetic order.  C is never read when beta == 0. */
void na_oracle_gemm_f64(size_t m, size_t k, size_t n, double alpha,
                        const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                        const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                        double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
/* The gemv/axcpy fallback alone (blas_uninit.rs:320-331, 127-177, 32-76), any size. */
void na_oracle_gemm_fallback_f64(size_t m, size_t k, size_t n, double alpha,
                                 const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                                 const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                                 double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
/* The matrixmultiply stand-in alone; nthreads > 1 parallelises the MC loop (the crate's optional
 * `threading` feature; nalgebra's default features leave it off, i.e. nthreads = 1). */
void na_oracle_dgemm_mm(size_t m, size_t k, size_t n, double alpha,
                        const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                        const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                        double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc, int nthreads);
void na_oracle_gemm_f32(size_t m, size_t k, size_t n, float alpha,
                        const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                        const float* b, ptrdiff_t rsb, ptrdiff_t csb,
                        float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc);
/* C <- alpha*A^T*B + beta*C, A is k x m: src/base/blas.rs:770-803 (gemm_tr), 503-540, 23-168. */
void na_oracle_gemm_tr_f64(size_t m, size_t k, size_t n, double alpha,
                           const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                           const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                           double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
/* dotx, 8-accumulator order: src/base/blas.rs:23-168. */
double na_oracle_dot_f64(size_t n, const double* x, ptrdiff_t incx, const double* y, ptrdiff_t incy);

/* ---- Cholesky: src/linalg/cholesky.rs:221-272 (new_internal) ---- */
/* Returns 0 on success (Some), 1 on failure (None); *fail_col = failing column. Lower only. */
int na_oracle_cholesky_f64(size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col);
/* solve_mut: cholesky.rs:122-129 -> solve.rs:488-519 + 697-755. */
void na_oracle_cholesky_solve_f64(size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs);

/* ---- LU: src/linalg/lu.rs:93-122, 337-389; min_max.rs:221-240; permutation_sequence.rs ---- */
/* swaps holds (i, i2) pairs, 2*min(m,n) entries of capacity; *nswaps = PermutationSequence::len. */
void na_oracle_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* swaps, size_t* nswaps);
/* lu.rs:242-260: returns 1 (true) on success, 0 (false) on an exactly-zero U diagonal. */
int na_oracle_lu_solve_f64(size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                           double* b, size_t ldb, size_t nrhs);
size_t na_oracle_icamax_f64(size_t n, const double* x, ptrdiff_t incx);
void na_oracle_permute_rows_f64(const size_t* swaps, size_t nswaps, double* b, size_t ldb, size_t ncols);
void na_oracle_inv_permute_rows_f64(const size_t* swaps, size_t nswaps, double* b, size_t ldb, size_t ncols);
double na_oracle_lu_determinant_f64(size_t n, const double* lu, size_t lda, size_t nswaps);
/* lu.rs:51-86 try_invert_to: returns 1 on success. `a` is consumed, out (n x n, ldo) receives the inverse. */
int na_oracle_try_invert_f64(size_t n, double* a, size_t lda, double* out, size_t ldo);

/* ---- QR: src/linalg/qr.rs:55-76; householder.rs:19-85; geometry/reflection.rs:70-83 ---- */
void na_oracle_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag);
/* qr.rs:108-129: q is m x min(m,n). */
void na_oracle_qr_q_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq);
/* qr.rs:81-89: r is min(m,n) x n. */
void na_oracle_qr_r_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* r, size_t ldr);
/* qr.rs:157-171. */
void na_oracle_qr_q_tr_mul_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag,
                               double* b, size_t ldb, size_t nrhs);
/* qr.rs:204-256 (square only): returns 1 on success, 0 on a zero diagonal. */
int na_oracle_qr_solve_f64(size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs);

/* ---- triangular solves: src/linalg/solve.rs:55-182 (checked; return 1/0) ---- */
int na_oracle_solve_lower_f64(size_t n, const double* a, size_t lda, double* b, size_t ldb, size_t nrhs);
int na_oracle_solve_upper_f64(size_t n, const double* a, size_t lda, double* b, size_t ldb, size_t nrhs);
int na_oracle_solve_lower_with_diag_f64(size_t n, const double* a, size_t lda, double diag, double* b, size_t ldb, size_t nrhs);

/* src/linalg/full_piv_lu.rs:56-91 (icamax_full: src/base/min_max.rs:146-167).  p_swaps / q_swaps: 2*min(m,n) entries each. */
void na_oracle_full_piv_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* p_swaps, size_t* np, size_t* q_swaps, size_t* nq);
/* src/linalg/col_piv_qr.rs:56-93.  Storage as na_oracle_qr_f64. */
void na_oracle_col_piv_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag, size_t* p_swaps, size_t* np);

/* ---- two-sided Householder reductions (householder.rs:61-127, reflection.rs:70-131) ---- */
/* src/linalg/hessenberg.rs:61-100: axes in column i rows i + 1.., H in the upper Hessenberg part, subdiag: n - 1 signed norms. */
void na_oracle_hessenberg_f64(size_t n, double* a, size_t lda, double* subdiag);
/* src/linalg/householder.rs:132-152 (assemble_q), used by Hessenberg::q and SymmetricTridiagonal::q. */
void na_oracle_assemble_q_f64(size_t n, const double* m, size_t lda, const double* signs, double* q, size_t ldq);
/* src/linalg/symmetric_tridiagonal.rs:54-95 (lower triangle only; blas.rs:359-420 xxgemv, 868-900 xxgerx). */
void na_oracle_symmetric_tridiagonal_f64(size_t n, double* a, size_t lda, double* off_diagonal);
/* src/linalg/bidiagonal.rs:74-150; returns upper_diagonal (m >= n). */
int na_oracle_bidiagonal_f64(size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal);
/* src/linalg/bidiagonal.rs:205-240 and 244-283. */
void na_oracle_bidiagonal_u_f64(size_t m, size_t n, const double* uv, size_t lda, const double* diagonal, const double* off_diagonal,
                                double* u, size_t ldu);
void na_oracle_bidiagonal_v_t_f64(size_t m, size_t n, const double* uv, size_t lda, const double* diagonal, const double* off_diagonal,
                                  double* vt, size_t ldvt);

#ifdef __cplusplus
}
#endif
#endif
