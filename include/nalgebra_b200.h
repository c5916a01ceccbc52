/*
 * nalgebra_b200.h -- C ABI of the B200-native dense back end for dimforge/nalgebra's hot path.
 *
 * One shared library (libnalgebra_b200.so, hand-written sm_100a CUDA) replaces, for DMatrix<f64>
 * (and the f32 GEMM), exactly the calls the reference makes on this path.  nalgebra has no plugin
 * API; the two seams a maintainer binds are:
 *
 *   seam 1  the single call into the third-party crate:
 *             matrixmultiply::dgemm / sgemm, /root/reference/src/base/blas_uninit.rs:298-313, 276-291
 *   seam 2  the back-end-crate pattern of nalgebra-lapack (same-named structs whose constructors
 *             call C symbols resolved at link time, nalgebra-lapack/src/lib.rs:33-36), but keeping
 *             CORE nalgebra's layouts: Cholesky{chol}, LU{lu,p}, PermutationSequence, QR{qr,diag}.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Column-major; strides and leading dimensions in ELEMENTS.
 *   - The caller owns every buffer; nothing is retained after return.  Factorizations are in place,
 *     like the reference (the Rust side passes the matrix it consumed by value).
 *   - Host-pointer entry points block until the result is visible to the host.  `_dev` twins take
 *     device pointers plus a cudaStream_t (as void*) and are asynchronous on that stream, except
 *     where a status must be returned (noted below).
 *   - Shape errors are panics on the Rust side before the FFI; the C side still returns NA_EINVAL.
 *     Numerical outcomes are values (NA_NOT_PD, NA_SINGULAR), never aborts.  CUDA failures map to
 *     NA_ECUDA and na_last_error() holds the message.
 *   - There is NO CPU fallback: without a usable sm_100 device every compute call fails with
 *     NA_ECUDA.
 *   - Thread-safe: calls may arrive concurrently from several host threads (the reference types
 *     are Send + Sync); host-pointer calls serialise on an internal lock, `_dev` calls only touch
 *     the stream they are given.
 */
#ifndef NALGEBRA_B200_H
#define NALGEBRA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define NAB_API __attribute__((visibility("default")))
#else
#define NAB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum na_status {
    NA_OK = 0,
    NA_NOT_PD = 1,     /* Cholesky::new -> None            (src/linalg/cholesky.rs:267-268) */
    NA_SINGULAR = 2,   /* solve_mut -> false               (src/linalg/solve.rs:169-171, qr.rs:242-244) */
    NA_EINVAL = -1,    /* shape/stride error (a panic in the reference: blas_uninit.rs:244-252) */
    NA_ECUDA = -2,     /* CUDA runtime/driver failure, or no sm_100 device */
    NA_ENOMEM = -3,    /* device or pinned-host allocation failed */
    NA_ENCCL = -4      /* reserved for the multi-GPU layer */
} na_status;

/* ---- context ------------------------------------------------------------------------------ */
/* Binds the calling process to `device` (default 0 when never called) and creates the internal
 * stream.  Idempotent. */
NAB_API int na_init(int device);
NAB_API int na_shutdown(void);
/* Message of the last failure on the calling thread ("" when none). */
NAB_API const char* na_last_error(void);
/* "nalgebra_b200 <version> sm_100a". */
NAB_API const char* na_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
NAB_API uint64_t na_kernel_launches(void);

/* ---- device / pinned memory helpers (so callers without a CUDA binding can use *_dev) ------ */
NAB_API int na_dev_malloc(void** ptr, size_t bytes);
NAB_API int na_dev_free(void* ptr);
NAB_API int na_host_alloc_pinned(void** ptr, size_t bytes);
NAB_API int na_host_free_pinned(void* ptr);
NAB_API int na_memcpy_h2d(void* dst, const void* src, size_t bytes);
NAB_API int na_memcpy_d2h(void* dst, const void* src, size_t bytes);
NAB_API int na_dev_synchronize(void);
/* Peer memory for the multi-GPU GEMM (one process per GPU): a buffer from na_dev_malloc is exported with
 * na_ipc_get_handle (64 opaque bytes, sent to the peer process by any means), mapped by the peer with
 * na_ipc_open_handle and read with na_memcpy_peer_async -- a copy-engine transfer over NVLink that takes no SM from
 * the running GEMM.  na_ipc_close_handle unmaps. */
NAB_API int na_ipc_get_handle(const void* dev_ptr, unsigned char handle[64]);
NAB_API int na_ipc_open_handle(const unsigned char handle[64], void** dev_ptr);
NAB_API int na_ipc_close_handle(void* dev_ptr);
NAB_API int na_memcpy_peer_async(void* dst, const void* src, size_t bytes, void* stream);
/* Fills a device matrix with the counter-based U[0,1) generator shared with the oracle:
 * a(i,j) = rand01(seed, i + j*nrows)  (the distribution of DMatrix::new_random,
 * src/base/construction.rs:293-299). */
NAB_API int na_fill_uniform_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed, void* stream);
/* The block [row0, row0+nrows) x [col0, col0+ncols) of the global_rows-tall matrix of the same
 * generator (sharded inputs: every rank fills only the panels it owns). */
NAB_API int na_fill_uniform_block_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                                      size_t row0, size_t col0, size_t global_rows, void* stream);

/* ---- seam 1: GEMM ------------------------------------------------------------------------- */
/* Replaces matrixmultiply::dgemm(m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc)
 * as called at src/base/blas_uninit.rs:298-313 (and through it Matrix::gemm blas.rs:729-746,
 * mul_to ops.rs:783-795, `&A * &B` ops.rs:554-574; gemm_tr/tr_mul are the same call with A's
 * strides swapped).  C is m x n, A is m x k, B is k x n, arbitrary element strides.
 * C is never read when beta == 0 (it may be uninitialised memory, ops.rs:567-571).
 * k == 0 scales C by beta (or zeroes it), blas_uninit.rs:258-269. */
NAB_API int na_dgemm(size_t m, size_t k, size_t n, double alpha,
             const double* a, ptrdiff_t rsa, ptrdiff_t csa,
             const double* b, ptrdiff_t rsb, ptrdiff_t csb,
             double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
/* Replaces matrixmultiply::sgemm, src/base/blas_uninit.rs:276-291. */
NAB_API int na_sgemm(size_t m, size_t k, size_t n, float alpha,
             const float* a, ptrdiff_t rsa, ptrdiff_t csa,
             const float* b, ptrdiff_t rsb, ptrdiff_t csb,
             float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc);
NAB_API int na_dgemm_dev(size_t m, size_t k, size_t n, double alpha,
                 const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                 double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc, void* stream);
NAB_API int na_sgemm_dev(size_t m, size_t k, size_t n, float alpha,
                 const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const float* b, ptrdiff_t rsb, ptrdiff_t csb,
                 float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc, void* stream);

/* C <- alpha * A * A^T + beta * C on the LOWER triangle (incl. diagonal) of the n x n column-major C only (A is
 * n x k, any strides); the strict upper triangle is neither read nor written.  The product the reference's SPD
 * recipes form with `&m * m.transpose()` (benches/linalg/cholesky.rs:3-11, src/debug/random_sdp.rs:34-45) and of
 * which Cholesky::new reads exactly this triangle; SURVEY.md 8(f)1. */
NAB_API int na_dsyrk_lower(size_t n, size_t k, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                           double beta, double* c, size_t ldc);
/* gemv_uninit (src/base/blas_uninit.rs:127-177; Matrix::gemv blas.rs:421-440; gemv_tr = the same call with A's
 * strides swapped, blas.rs:503-540): y (m entries, stride incy) <- alpha * A (m x n) * x + beta * y.  y is not read
 * when beta == 0; n == 0 scales or zeroes y (:152-160).  The fallback of gemm_uninit for small / non-Dyn shapes. */
NAB_API int na_dgemv(size_t m, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                     const double* x, ptrdiff_t incx, double beta, double* y, ptrdiff_t incy);
NAB_API int na_dgemv_dev(size_t m, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                         const double* x, ptrdiff_t incx, double beta, double* y, ptrdiff_t incy, void* stream);
/* axcpy (src/base/blas_uninit.rs:86-117): y <- a * x * c + b * y with the reference's rounding order ((a*x)*c, unfused);
 * y is not read when b == 0.  Device pointers. */
NAB_API int na_daxcpy_dev(size_t n, double a, const double* x, ptrdiff_t incx, double c, double b, double* y, ptrdiff_t incy,
                          void* stream);

/* ---- seam 2: Cholesky --------------------------------------------------------------------- */
/* Cholesky::new / new_with_substitute (src/linalg/cholesky.rs:196-272).  On NA_OK the lower
 * triangle (incl. diagonal) of `a` holds L; the strict upper triangle is never read or written
 * (tests/linalg/cholesky.rs:3-12).  NA_NOT_PD == `None`; *fail_col (may be NULL) receives the
 * first column whose pivot was <= 0 or NaN.  use_sub/sub = new_with_substitute's `substitute`.
 * One deliberate deviation from the reference's arithmetic: a column is scaled by the reciprocal of
 * sqrt(pivot) where cholesky.rs:261-262 divides (<= 1 ulp per entry; eight serial FP64 divisions per
 * thread and column would lengthen the latency-bound 128-column leaf by a quarter). */
NAB_API int na_cholesky_f64(size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col);
/* Synchronises `stream` before returning (the status is a value). */
NAB_API int na_cholesky_f64_dev(size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col, void* stream);
/* The same without the status read-back (no host synchronisation for n <= 1024): the first failing column, plus
 * col_offset, is atomicMin'ed into the DEVICE word *fail_col_dev, which the caller set to UINT64_MAX beforehand
 * (several panels of a block-cyclic factorization may share one word; it stays UINT64_MAX on success). */
NAB_API int na_cholesky_f64_dev_async(size_t n, double* a, size_t lda, int use_sub, double sub, uint64_t* fail_col_dev,
                                      size_t col_offset, void* stream);
/* Cholesky::solve_mut (cholesky.rs:122-129): b <- (L L^T)^-1 b, b is n x nrhs. */
NAB_API int na_cholesky_solve_f64(size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs);
NAB_API int na_cholesky_solve_f64_dev(size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs, void* stream);

/* ---- seam 2: LU with partial pivoting ------------------------------------------------------ */
/* LU::new (src/linalg/lu.rs:93-122).  On return `a` holds the packed factors (strict lower = L
 * multipliers, upper incl. diagonal = U; row swaps applied to whole rows, LAPACK getrf layout)
 * and swaps[2*s], swaps[2*s+1] (s < *nswaps) is PermutationSequence's s-th pair (i, i2)
 * (src/linalg/permutation_sequence.rs:28-34, 84-93: only non-trivial swaps are stored, in
 * application order).  `swaps` has room for 2*min(m,n) entries; unused entries are set to 0.
 * Pivot rule = icamax (src/base/min_max.rs:221-240): largest |x|, lowest index wins ties; an
 * all-zero column is skipped without a swap (lu.rs:107-110).  Never fails numerically. */
NAB_API int na_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* swaps, size_t* nswaps);
/* swaps/nswaps are HOST pointers; synchronises `stream` before returning. */
NAB_API int na_lu_f64_dev(size_t m, size_t n, double* a, size_t lda, size_t* swaps, size_t* nswaps, void* stream);
/* The same without the pivot read-back: ipiv_dev (DEVICE, min(m, n) int32) receives the 0-based pivot row of every
 * column (LAPACK's ipiv minus one; ipiv[i] == i where the reference records no swap).  Asynchronous on `stream`. */
NAB_API int na_lu_f64_dev_async(size_t m, size_t n, double* a, size_t lda, int32_t* ipiv_dev, void* stream);
/* Applies the row interchanges (row0 + s) <-> (row0 + ipiv_dev[s]), s = 0 .. k-1 in order, to the nrows x ncols
 * column-major device matrix `a` (PermutationSequence::permute_rows for a pivot vector that never left the device). */
NAB_API int na_apply_ipiv_f64_dev(size_t nrows, double* a, size_t lda, size_t ncols, const int32_t* ipiv_dev, size_t k,
                                  size_t row0, void* stream);
/* LU::solve_mut (lu.rs:242-260): permute rows of b, unit-lower solve, upper solve.
 * NA_SINGULAR == `false` (an exactly-zero U[i,i]); b is then garbage, as in the reference. */
NAB_API int na_lu_solve_f64(size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                    double* b, size_t ldb, size_t nrhs);
NAB_API int na_lu_solve_f64_dev(size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                        double* b, size_t ldb, size_t nrhs, void* stream);

/* ---- seam 2: Householder QR ---------------------------------------------------------------- */
/* QR::new (src/linalg/qr.rs:55-76).  nalgebra's storage, NOT LAPACK's: column i, rows i..m of `a`
 * hold the unit-2-norm Householder axis u_i (first component included); the strict upper triangle
 * holds R's off-diagonal; diag[i] (min(m,n) entries) is the signed value returned by
 * reflection_axis_mut (householder.rs:19-53), so R[i,i] = |diag[i]| (qr.rs:87). */
NAB_API int na_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag);
NAB_API int na_qr_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diag, void* stream);
/* QR::q (qr.rs:108-129): q is m x min(m,n). */
NAB_API int na_qr_q_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq);
NAB_API int na_qr_q_f64_dev(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq, void* stream);
/* QR::q_tr_mul (qr.rs:157-171): b <- Q^T b, b is m x nrhs. */
NAB_API int na_qr_q_tr_mul_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag,
                       double* b, size_t ldb, size_t nrhs);
NAB_API int na_qr_q_tr_mul_f64_dev(size_t m, size_t n, const double* qr, size_t lda, const double* diag,
                           double* b, size_t ldb, size_t nrhs, void* stream);
/* QR::solve_mut (qr.rs:204-256), square only.  NA_SINGULAR == `false` (a zero in diag). */
NAB_API int na_qr_solve_f64(size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs);

/* ---- seam 2: the factorizations that pivot on the largest entry of the trailing matrix ------------ */
/* FullPivLU::new (src/linalg/full_piv_lu.rs:56-91): P A Q = L U.  Pivot of step i = icamax_full of the trailing
 * matrix (src/base/min_max.rs:146-167: column-major scan, strict >); whole columns and whole rows are swapped, the
 * elimination is lu::gauss_step(_swap) (IEEE reciprocal, unfused multiply-add), so `a` and both sequences are
 * bit-identical to the reference's.  An exactly zero pivot stops the factorization (:73-76).  p_swaps / q_swaps: room for
 * 2*min(m,n) entries each, PermutationSequence pairs (i, i2) with i != i2 only, unused entries 0. */
NAB_API int na_full_piv_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* p_swaps, size_t* np, size_t* q_swaps, size_t* nq);
/* a: DEVICE; the four outputs are HOST pointers; synchronises `stream` before returning. */
NAB_API int na_full_piv_lu_f64_dev(size_t m, size_t n, double* a, size_t lda, size_t* p_swaps, size_t* np, size_t* q_swaps, size_t* nq,
                                   void* stream);
/* ColPivQR::new (src/linalg/col_piv_qr.rs:56-93): A P = Q R, the pivot column of step i is the column of icamax_full of the
 * trailing matrix; storage as na_qr_f64 (unit Householder axes below the diagonal, diag = signed norms), so na_qr_q_f64 /
 * na_qr_q_tr_mul_f64 apply to it unchanged.  p_swaps: room for 2*min(m,n) entries. */
NAB_API int na_col_piv_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag, size_t* p_swaps, size_t* np);
/* a, diag: DEVICE; p_swaps / np: HOST; synchronises `stream` before returning. */
NAB_API int na_col_piv_qr_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diag, size_t* p_swaps, size_t* np, void* stream);

/* ---- two-sided Householder reductions (src/linalg/householder.rs:61-127) ---------------------- */
/* Hessenberg::new (src/linalg/hessenberg.rs:61-100): a (n x n) is overwritten with nalgebra's packed `hess`: H in the upper
 * Hessenberg part except its first subdiagonal, the unit Householder axis of step i in column i, rows i + 1..; subdiag[i]
 * (n - 1 entries) = the signed norm clear_column_unchecked returns (H[i + 1, i] = |subdiag[i]|, the sign feeds
 * householder::assemble_q).  n == 0 -> NA_EINVAL (the reference panics). */
NAB_API int na_hessenberg_f64(size_t n, double* a, size_t lda, double* subdiag);
/* a, subdiag: DEVICE.  Asynchronous on `stream`. */
NAB_API int na_hessenberg_f64_dev(size_t n, double* a, size_t lda, double* subdiag, void* stream);
/* SymmetricTridiagonal::new (src/linalg/symmetric_tridiagonal.rs:54-95): only the lower triangle of a is read and written;
 * on return its diagonal is the diagonal of T, column i rows i + 1.. the axis of step i, off_diagonal[i] (n - 1 entries) the
 * signed norm (T[i + 1, i] = |off_diagonal[i]|). */
NAB_API int na_symmetric_tridiagonal_f64(size_t n, double* a, size_t lda, double* off_diagonal);
NAB_API int na_symmetric_tridiagonal_f64_dev(size_t n, double* a, size_t lda, double* off_diagonal, void* stream);
/* Bidiagonal::new (src/linalg/bidiagonal.rs:74-150): a (m x n) is overwritten with nalgebra's packed `uv` (m >= n: column
 * axes in column i rows i.., row axes in row i columns i + 1..; m < n: row axes in row i columns i.., column axes in column i
 * rows i + 1..); diagonal: min(m, n) signed norms, off_diagonal: min(m, n) - 1.  upper_diagonal = (m >= n). */
NAB_API int na_bidiagonal_f64(size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal);
NAB_API int na_bidiagonal_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal, void* stream);

/* ---- triangular solves (src/linalg/solve.rs:55-182) ---------------------------------------- */
/* op(T) x = b in place on b (n x nrhs).  lower: 1 = lower, 0 = upper triangle of `t` is used.
 * trans: 0 = T, 1 = T^T.  unit_diag: 1 = implicit unit diagonal (solve_lower_triangular_with_diag_mut
 * with diag = 1, solve.rs:106-133).  NA_SINGULAR on an exactly-zero diagonal entry. */
NAB_API int na_tri_solve_f64(int lower, int trans, int unit_diag, size_t n, const double* t, size_t ldt,
                     double* b, size_t ldb, size_t nrhs);
NAB_API int na_tri_solve_f64_dev(int lower, int trans, int unit_diag, size_t n, const double* t, size_t ldt,
                         double* b, size_t ldb, size_t nrhs, void* stream);

/* Upper bound on the CTAs (= SMs: the GEMM kernels are persistent, one CTA per SM) that GEMM launches
 * issued by the CALLING THREAD may use; 0 restores "all SMs".  Lets a caller keep SMs free for work on
 * other streams (NCCL copy kernels while panels are staged over NVLink, a concurrent panel kernel).
 * The setting is thread-local (it has no effect on launches made from other host threads) and stays in
 * force across factorization calls: the blocked drivers combine it with their own panel/bulk split
 * (the smaller limit applies) and never clear it. */
NAB_API int na_set_gemm_sm_limit(int max_ctas);

/* Tuning / diagnostic switches, never needed for correctness.  Keys: "lu_lookahead" (0: factor with the plain
 * recursive driver instead of the two-stream look-ahead driver; the tests compare the two paths' pivots);
 * "qr_reg_leaf" (0: shared-memory GEQR2 leaf + two-stream look-ahead driver instead of the register-resident leaf +
 * plain outer loop); "qr_fused" (0: in-panel block reflectors as GEMM sequences instead of the fused kernel); "ts_fused" (Hessenberg /
 * SymmetricTridiagonal: 1 = one fused pass per step, 0 = product pass + update pass, -1 = by size, the default). */
NAB_API int na_set_tuning(const char* key, long value);

/* ---- building blocks of the multi-GPU (1D block-cyclic) factorizations, device pointers ------ */
/* General triangular solve with many right-hand sides, in place on B (m x n, ldb):
 *   side_right = 0:  op(T) X = B  (T is m x m)      side_right = 1:  X op(T) = B  (T is n x n)
 * lower / trans / unit_diag as in na_tri_solve_f64.  No singularity check (the blocked drivers'
 * *_unchecked use, src/linalg/solve.rs:488-580). */
NAB_API int na_trsm_f64_dev(int side_right, int lower, int trans, int unit_diag, size_t m, size_t n,
                            const double* t, size_t ldt, double* b, size_t ldb, void* stream);
/* PermutationSequence::permute_rows / inv_permute_rows (src/linalg/permutation_sequence.rs:97-116) on a
 * device matrix: applies the swaps (swaps[2s], swaps[2s+1]), s < nswaps (HOST array, application order;
 * inverse != 0 applies them in reverse) to the rows of the nrows x ncols column-major matrix `a`. */
NAB_API int na_permute_rows_f64_dev(size_t nrows, double* a, size_t lda, size_t ncols,
                                    const size_t* swaps, size_t nswaps, int inverse, void* stream);
/* C <- alpha*A*B + beta*C restricted to the lower trapezoid (row >= col) of the m x n (m >= n)
 * column-major C: the SYRK-shaped trailing update of Cholesky (cholesky.rs:226-235 never touches the
 * strict upper triangle, and neither does this). */
NAB_API int na_dgemm_lower_dev(size_t m, size_t k, size_t n, double alpha,
                               const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                               const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                               double beta, double* c, size_t ldc, void* stream);
/* The block [row0, +nrows) x [col0, +ncols) of the n x n SPD test matrix (B + B^T)/2 + n*I,
 * B(i,j) = rand01(seed, i + j*n)  (Cfg 3 (ii)). */
NAB_API int na_fill_spd_block_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                                  size_t row0, size_t col0, size_t n, void* stream);

/* ---- seam 2': LAPACK symbols (Fortran ABI) for `nalgebra-lapack --features lapack-custom` -------------------------
 * /root/reference/nalgebra-lapack/src/lib.rs:33-36; call sites cholesky.rs:181-224, lu.rs:351-446, qr.rs:166-237 and
 * 369-590.  Every argument by pointer, INTEGER = int32, characters as one byte, info through the last argument; HOST
 * pointers, LAPACK layouts (potrf 'L'/'U' factor in place, getrf L\U + 1-based ipiv, geqrf R + reflector vectors + tau).
 * Workspace queries (lwork = -1) answer 1.  Without an sm_100 device info = -1000 (no CPU fallback). */
NAB_API void dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
NAB_API void dpotrs_(const char* uplo, const int* n, const int* nrhs, const double* a, const int* lda, double* b, const int* ldb, int* info);
NAB_API void dpotri_(const char* uplo, const int* n, double* a, const int* lda, int* info);
NAB_API void dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
NAB_API void dlaswp_(const int* n, double* a, const int* lda, const int* k1, const int* k2, const int* ipiv, const int* incx);
NAB_API void dgetrs_(const char* trans, const int* n, const int* nrhs, const double* a, const int* lda, const int* ipiv, double* b,
                     const int* ldb, int* info);
NAB_API void dgetri_(const int* n, double* a, const int* lda, const int* ipiv, double* work, const int* lwork, int* info);
NAB_API void dgeqrf_(const int* m, const int* n, double* a, const int* lda, double* tau, double* work, const int* lwork, int* info);
NAB_API void dormqr_(const char* side, const char* trans, const int* m, const int* n, const int* k, const double* a, const int* lda,
                     const double* tau, double* c, const int* ldc, double* work, const int* lwork, int* info);
NAB_API void dorgqr_(const int* m, const int* n, const int* k, double* a, const int* lda, const double* tau, double* work,
                     const int* lwork, int* info);
NAB_API void dtrtrs_(const char* uplo, const char* trans, const char* diag, const int* n, const int* nrhs, const double* a, const int* lda,
                     double* b, const int* ldb, int* info);

#ifdef __cplusplus
}
#endif
#endif /* NALGEBRA_B200_H */
