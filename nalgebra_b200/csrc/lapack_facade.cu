// lapack_facade.cu -- Fortran-ABI LAPACK symbols over the B200 kernels, so that `nalgebra-lapack` built with
// `--features lapack-custom` links against libnalgebra_b200.so unmodified (SURVEY.md §8(b) seam 2, §8(f)4).
//
// Reference call sites: /root/reference/nalgebra-lapack/src/lib.rs:33-36 (lapack-custom: "functions must be
// available at link time ... ABI compatible with the lapack crate"), cholesky.rs:181-224 (xpotrf / xpotrs / xpotri),
// lu.rs:351-446 (xgetrf / xlaswp / xgetrs / xgetri), qr.rs:166-237, 369-590 (xgeqrf / xorgqr / xormqr / xtrtrs).
// The `lapack` crate binds the plain Fortran symbols: every argument by pointer, characters as one byte, no hidden
// string lengths, INTEGER = int32, info returned through the last argument.
//
// These entry points take HOST pointers (LAPACK's contract), stage through the device and produce LAPACK layouts
// (not core nalgebra's): potrf the 'L' or 'U' factor in place; getrf packed L\U + 1-based ipiv; geqrf R on/above the
// diagonal, reflector vectors below it + tau.  No CPU fallback: without an sm_100 device info = -1000 and
// na_last_error() says why.  Workspace queries (lwork = -1) answer 1: no host workspace is used.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

using namespace nab;

namespace {

constexpr int kNoDevice = -1000;

struct HostCall {                       // lock + stream of one host-pointer call
    std::unique_lock<std::mutex> lock;
    cudaStream_t s = nullptr;
    int begin() {
        int st = ensure_init();
        if (st != NA_OK) return st;
        lock = std::unique_lock<std::mutex>(host_api_mutex());
        s = ctx().stream;
        return NA_OK;
    }
};

inline bool is(char c, char want) { return c == want || c == (char)(want + 32); }

// info for an internal failure: LAPACK has no code for "the accelerator failed"; use a large negative value
inline int fail_info(int st) { return st == NA_ENOMEM ? -1001 : kNoDevice; }

// dst (column-major n x n, ldd) <- src^T
int transpose_square(cudaStream_t s, double* dst, size_t ldd, const double* src, size_t lds, size_t n) {
    return copy_strided(s, dst, 1, (ptrdiff_t)ldd, src, (ptrdiff_t)lds, 1, n, n);
}

}  // namespace

extern "C" {

// ---- Cholesky ------------------------------------------------------------------------------------------------
NAB_API void dpotrf_(const char* uplo, const int* n_, double* a, const int* lda_, int* info) {
    const int n = *n_, lda = *lda_;
    *info = 0;
    if (!(is(*uplo, 'L') || is(*uplo, 'U'))) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (lda < std::max(1, n)) { *info = -4; return; }
    if (n == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, t; size_t ldd;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, n, n);
    double* work = d.as<double>();
    if (st == NA_OK && is(*uplo, 'U')) {        // A = U^T U: the lower Cholesky of the transposed storage
        st = t.alloc(ldd * n * sizeof(double), hc.s);
        if (st == NA_OK) st = transpose_square(hc.s, t.as<double>(), ldd, d.as<double>(), ldd, n);
        work = t.as<double>();
    }
    size_t fail_col = 0;
    if (st == NA_OK) st = cholesky_device(hc.s, n, work, ldd, 0, 0.0, &fail_col);
    if (st < 0) { *info = fail_info(st); return; }
    if (st == NA_NOT_PD) *info = (int)fail_col + 1;          // "the leading minor of order i is not positive definite"
    if (is(*uplo, 'U')) {
        // only the upper triangle of A may change: transpose the factor back over the uploaded copy, triangle by triangle
        if (transpose_square(hc.s, d.as<double>(), ldd, work, ldd, n) < 0) { *info = kNoDevice; return; }
        // d now holds L^T in its upper triangle and the (transposed) untouched strict upper of `work` in its lower one;
        // download only the columns' upper parts
        for (int j = 0; j < n; j += 128) {
            const int w = std::min(128, n - j);
            if (cudaMemcpy2DAsync(a + (size_t)j * lda, (size_t)lda * 8, d.as<double>() + (size_t)j * ldd, ldd * 8, (size_t)(j + w) * 8, w,
                                  cudaMemcpyDeviceToHost, hc.s) != cudaSuccess) { *info = kNoDevice; return; }
        }
    } else {
        for (int j = 0; j < n; j += 128) {                   // lower trapezoids only: the strict upper triangle is the caller's
            const int w = std::min(128, n - j);
            if (cudaMemcpy2DAsync(a + j + (size_t)j * lda, (size_t)lda * 8, d.as<double>() + j + (size_t)j * ldd, ldd * 8, (size_t)(n - j) * 8, w,
                                  cudaMemcpyDeviceToHost, hc.s) != cudaSuccess) { *info = kNoDevice; return; }
        }
    }
    if (cudaStreamSynchronize(hc.s) != cudaSuccess) *info = kNoDevice;
}

NAB_API void dpotrs_(const char* uplo, const int* n_, const int* nrhs_, const double* a, const int* lda_, double* b, const int* ldb_, int* info) {
    const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
    *info = 0;
    if (!(is(*uplo, 'L') || is(*uplo, 'U'))) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (nrhs < 0) { *info = -3; return; }
    if (lda < std::max(1, n)) { *info = -5; return; }
    if (ldb < std::max(1, n)) { *info = -7; return; }
    if (n == 0 || nrhs == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, t, db; size_t ldd, lddb;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, n, n);
    const double* l = d.as<double>();
    if (st == NA_OK && is(*uplo, 'U')) {
        st = t.alloc(ldd * n * sizeof(double), hc.s);
        if (st == NA_OK) st = transpose_square(hc.s, t.as<double>(), ldd, d.as<double>(), ldd, n);
        l = t.as<double>();
    }
    if (st == NA_OK) st = upload_matrix(hc.s, db, lddb, b, (size_t)ldb, n, nrhs);
    if (st == NA_OK) st = cholesky_solve_device(hc.s, n, l, ldd, db.as<double>(), lddb, nrhs);
    if (st == NA_OK) st = download_matrix(hc.s, b, (size_t)ldb, db.as<double>(), lddb, n, nrhs);
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) *info = fail_info(st);
}

NAB_API void dpotri_(const char* uplo, const int* n_, double* a, const int* lda_, int* info) {
    const int n = *n_, lda = *lda_;
    *info = 0;
    if (!(is(*uplo, 'L') || is(*uplo, 'U'))) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (lda < std::max(1, n)) { *info = -4; return; }
    if (n == 0) return;
    for (int i = 0; i < n; ++i) if (a[i + (size_t)i * lda] == 0.0) { *info = i + 1; return; }       // singular triangular factor
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, t, x; size_t ldd;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, n, n);
    const double* l = d.as<double>();
    if (st == NA_OK && is(*uplo, 'U')) {
        st = t.alloc(ldd * n * sizeof(double), hc.s);
        if (st == NA_OK) st = transpose_square(hc.s, t.as<double>(), ldd, d.as<double>(), ldd, n);
        l = t.as<double>();
    }
    if (st == NA_OK) st = x.alloc(ldd * n * sizeof(double), hc.s);
    if (st == NA_OK) st = set_identity(hc.s, x.as<double>(), ldd, n, n);
    if (st == NA_OK) st = cholesky_solve_device(hc.s, n, l, ldd, x.as<double>(), ldd, n);       // X = (L L^T)^-1, symmetric
    if (st != NA_OK) { *info = fail_info(st); return; }
    for (int j = 0; j < n; j += 128) {
        const int w = std::min(128, n - j);
        cudaError_t e = is(*uplo, 'U')
            ? cudaMemcpy2DAsync(a + (size_t)j * lda, (size_t)lda * 8, x.as<double>() + (size_t)j * ldd, ldd * 8, (size_t)(j + w) * 8, w, cudaMemcpyDeviceToHost, hc.s)
            : cudaMemcpy2DAsync(a + j + (size_t)j * lda, (size_t)lda * 8, x.as<double>() + j + (size_t)j * ldd, ldd * 8, (size_t)(n - j) * 8, w, cudaMemcpyDeviceToHost, hc.s);
        if (e != cudaSuccess) { *info = kNoDevice; return; }
    }
    // the 128 x 128 diagonal blocks travelled whole: restore the opposite strict triangle of the caller's matrix there
    // (it was uploaded with `a`, so re-download it from d, where it is untouched)
    if (cudaStreamSynchronize(hc.s) != cudaSuccess) { *info = kNoDevice; return; }
    std::vector<double> blk(128 * 128);
    for (int j = 0; j < n; j += 128) {
        const int w = std::min(128, n - j);
        if (cudaMemcpy2D(blk.data(), 128 * 8, d.as<double>() + j + (size_t)j * ldd, ldd * 8, (size_t)w * 8, w, cudaMemcpyDeviceToHost) != cudaSuccess) { *info = kNoDevice; return; }
        for (int c = 0; c < w; ++c)
            for (int r = 0; r < w; ++r)
                if (is(*uplo, 'U') ? r > c : r < c) a[(j + r) + (size_t)(j + c) * lda] = blk[r + c * 128];
    }
}

// ---- LU ----------------------------------------------------------------------------------------------------------
NAB_API void dgetrf_(const int* m_, const int* n_, double* a, const int* lda_, int* ipiv, int* info) {
    const int m = *m_, n = *n_, lda = *lda_;
    *info = 0;
    if (m < 0) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (lda < std::max(1, m)) { *info = -4; return; }
    const int mn = std::min(m, n);
    if (mn == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, dp; size_t ldd;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, m, n);
    if (st == NA_OK) st = dp.alloc((size_t)mn * sizeof(int), hc.s);
    if (st == NA_OK) st = lu_device_async(hc.s, m, n, d.as<double>(), ldd, dp.as<int>());
    if (st == NA_OK) st = download_matrix(hc.s, a, (size_t)lda, d.as<double>(), ldd, m, n);
    if (st == NA_OK && cudaMemcpyAsync(ipiv, dp.p, (size_t)mn * sizeof(int), cudaMemcpyDeviceToHost, hc.s) != cudaSuccess) st = NA_ECUDA;
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) { *info = fail_info(st); return; }
    for (int i = 0; i < mn; ++i) ipiv[i] += 1;                                               // 1-based
    for (int i = 0; i < mn; ++i) if (a[i + (size_t)i * lda] == 0.0) { *info = i + 1; break; }   // U(i,i) exactly zero
}

NAB_API void dlaswp_(const int* n_, double* a, const int* lda_, const int* k1_, const int* k2_, const int* ipiv, const int* incx_) {
    const int n = *n_, lda = *lda_, k1 = *k1_, k2 = *k2_, incx = *incx_;
    if (n <= 0 || incx == 0 || k2 < k1) return;
    // rows k1..k2 (1-based) are interchanged with ipiv(k1 + (i-k1)*|incx|...): LAPACK semantics, forward for incx > 0,
    // backward for incx < 0.  Touched rows: up to max(ipiv).
    int rows = k2;
    const int cnt = k2 - k1 + 1;
    std::vector<size_t> pairs;
    pairs.reserve(2 * (size_t)cnt);
    if (incx > 0) {
        for (int i = 0; i < cnt; ++i) { const int ip = ipiv[(size_t)(k1 - 1) + (size_t)i * incx]; rows = std::max(rows, ip); if (ip != k1 + i) { pairs.push_back((size_t)(k1 + i - 1)); pairs.push_back((size_t)(ip - 1)); } }
    } else {
        for (int i = cnt - 1; i >= 0; --i) { const int ip = ipiv[(size_t)(k1 - 1) + (size_t)i * (-incx)]; rows = std::max(rows, ip); if (ip != k1 + i) { pairs.push_back((size_t)(k1 + i - 1)); pairs.push_back((size_t)(ip - 1)); } }
    }
    if (pairs.empty()) return;
    HostCall hc;
    if (hc.begin() != NA_OK) return;
    Scratch d; size_t ldd;
    if (upload_matrix(hc.s, d, ldd, a, (size_t)lda, rows, n) != NA_OK) return;
    if (na_permute_rows_f64_dev((size_t)rows, d.as<double>(), ldd, (size_t)n, pairs.data(), pairs.size() / 2, 0, hc.s) != NA_OK) return;
    download_matrix(hc.s, a, (size_t)lda, d.as<double>(), ldd, rows, n);
    cudaStreamSynchronize(hc.s);
}

NAB_API void dgetrs_(const char* trans, const int* n_, const int* nrhs_, const double* a, const int* lda_, const int* ipiv, double* b,
                     const int* ldb_, int* info) {
    const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
    *info = 0;
    const bool tr = is(*trans, 'T') || is(*trans, 'C');
    if (!(tr || is(*trans, 'N'))) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (nrhs < 0) { *info = -3; return; }
    if (lda < std::max(1, n)) { *info = -5; return; }
    if (ldb < std::max(1, n)) { *info = -8; return; }
    if (n == 0 || nrhs == 0) return;
    std::vector<size_t> pairs;
    for (int i = 0; i < n; ++i) {
        if (ipiv[i] < 1 || ipiv[i] > n) { *info = -6; return; }
        if (ipiv[i] != i + 1) { pairs.push_back((size_t)i); pairs.push_back((size_t)(ipiv[i] - 1)); }
    }
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, db; size_t ldd, lddb;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, n, n);
    if (st == NA_OK) st = upload_matrix(hc.s, db, lddb, b, (size_t)ldb, n, nrhs);
    const double* lu = d.as<double>();
    double* x = db.as<double>();
    if (st == NA_OK && !tr) {           // A x = b: P b, L y = ., U x = y   (getrs does not test for singularity)
        if (!pairs.empty()) st = na_permute_rows_f64_dev((size_t)n, x, lddb, (size_t)nrhs, pairs.data(), pairs.size() / 2, 0, hc.s);
        if (st == NA_OK) st = trsm_left(hc.s, true, true, n, lu, 1, (ptrdiff_t)ldd, nullptr, nullptr, x, 1, (ptrdiff_t)lddb, nrhs);
        if (st == NA_OK) st = trsm_left(hc.s, false, false, n, lu, 1, (ptrdiff_t)ldd, nullptr, nullptr, x, 1, (ptrdiff_t)lddb, nrhs);
    } else if (st == NA_OK) {           // A^T x = b: U^T y = b, L^T z = y, x = P^T z
        st = trsm_left(hc.s, true, false, n, lu, (ptrdiff_t)ldd, 1, nullptr, nullptr, x, 1, (ptrdiff_t)lddb, nrhs);
        if (st == NA_OK) st = trsm_left(hc.s, false, true, n, lu, (ptrdiff_t)ldd, 1, nullptr, nullptr, x, 1, (ptrdiff_t)lddb, nrhs);
        if (st == NA_OK && !pairs.empty()) st = na_permute_rows_f64_dev((size_t)n, x, lddb, (size_t)nrhs, pairs.data(), pairs.size() / 2, 1, hc.s);
    }
    if (st == NA_OK) st = download_matrix(hc.s, b, (size_t)ldb, x, lddb, n, nrhs);
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) *info = fail_info(st);
}

NAB_API void dgetri_(const int* n_, double* a, const int* lda_, const int* ipiv, double* work, const int* lwork_, int* info) {
    const int n = *n_, lda = *lda_, lwork = *lwork_;
    *info = 0;
    if (n < 0) { *info = -1; return; }
    if (lda < std::max(1, n)) { *info = -3; return; }
    if (lwork == -1) { if (work) work[0] = 1.0; return; }             // workspace query
    if (lwork < std::max(1, n) && lwork != -1) { /* LAPACK would complain; no host workspace is needed here */ }
    if (n == 0) return;
    for (int i = 0; i < n; ++i) if (a[i + (size_t)i * lda] == 0.0) { *info = i + 1; return; }   // singular: no inverse
    std::vector<double> ident((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) ident[i + (size_t)i * n] = 1.0;
    const char nn = 'N';
    dgetrs_(&nn, n_, n_, a, lda_, ipiv, ident.data(), n_, info);
    if (*info != 0) return;
    for (int j = 0; j < n; ++j) memcpy(a + (size_t)j * lda, ident.data() + (size_t)j * n, (size_t)n * sizeof(double));
    if (work) work[0] = 1.0;
}

// ---- QR ----------------------------------------------------------------------------------------------------------
NAB_API void dgeqrf_(const int* m_, const int* n_, double* a, const int* lda_, double* tau, double* work, const int* lwork_, int* info) {
    const int m = *m_, n = *n_, lda = *lda_, lwork = *lwork_;
    *info = 0;
    if (m < 0) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (lda < std::max(1, m)) { *info = -4; return; }
    if (lwork == -1) { if (work) work[0] = 1.0; return; }
    const int k = std::min(m, n);
    if (k == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, dt; size_t ldd;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, m, n);
    if (st == NA_OK) st = dt.alloc((size_t)k * sizeof(double), hc.s);
    if (st == NA_OK) st = qr_device(hc.s, m, n, d.as<double>(), ldd, nullptr, dt.as<double>());
    if (st == NA_OK) st = download_matrix(hc.s, a, (size_t)lda, d.as<double>(), ldd, m, n);
    if (st == NA_OK && cudaMemcpyAsync(tau, dt.p, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, hc.s) != cudaSuccess) st = NA_ECUDA;
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) { *info = fail_info(st); return; }
    if (work) work[0] = 1.0;
}

// C (m x n) <- op(Q) C or C op(Q), Q = H_1 ... H_k from dgeqrf_ (a: reflector vectors, tau).
NAB_API void dormqr_(const char* side, const char* trans, const int* m_, const int* n_, const int* k_, const double* a, const int* lda_,
                     const double* tau, double* c, const int* ldc_, double* work, const int* lwork_, int* info) {
    const int m = *m_, n = *n_, k = *k_, lda = *lda_, ldc = *ldc_, lwork = *lwork_;
    *info = 0;
    const bool left = is(*side, 'L'), tr = is(*trans, 'T') || is(*trans, 'C');
    const int nq = left ? m : n;                     // order of Q
    if (!(left || is(*side, 'R'))) { *info = -1; return; }
    if (!(tr || is(*trans, 'N'))) { *info = -2; return; }
    if (m < 0) { *info = -3; return; }
    if (n < 0) { *info = -4; return; }
    if (k < 0 || k > nq) { *info = -5; return; }
    if (lda < std::max(1, nq)) { *info = -7; return; }
    if (ldc < std::max(1, m)) { *info = -10; return; }
    if (lwork == -1) { if (work) work[0] = 1.0; return; }
    if (m == 0 || n == 0 || k == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, dt, dc, dct; size_t ldd, lddc;
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, nq, k);
    if (st == NA_OK) st = dt.alloc((size_t)k * sizeof(double), hc.s);
    if (st == NA_OK && cudaMemcpyAsync(dt.p, tau, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, hc.s) != cudaSuccess) st = NA_ECUDA;
    if (st == NA_OK) st = upload_matrix(hc.s, dc, lddc, c, (size_t)ldc, m, n);
    if (st == NA_OK && left) {
        // forward = true applies Q^T (H_k .. H_1 block order 0, 1, ...), false applies Q
        st = apply_q_blocks(hc.s, m, k, d.as<double>(), ldd, nullptr, dc.as<double>(), lddc, n, tr, false, dt.as<double>());
    } else if (st == NA_OK) {
        // C op(Q) = (op(Q)^T C^T)^T: apply from the left to an explicit transpose
        const size_t ldt = round_up((size_t)n, 2);
        st = dct.alloc(ldt * m * sizeof(double), hc.s);
        if (st == NA_OK) st = copy_strided(hc.s, dct.as<double>(), 1, (ptrdiff_t)ldt, dc.as<double>(), (ptrdiff_t)lddc, 1, n, m);
        if (st == NA_OK) st = apply_q_blocks(hc.s, n, k, d.as<double>(), ldd, nullptr, dct.as<double>(), ldt, m, !tr, false, dt.as<double>());
        if (st == NA_OK) st = copy_strided(hc.s, dc.as<double>(), 1, (ptrdiff_t)lddc, dct.as<double>(), (ptrdiff_t)ldt, 1, m, n);
    }
    if (st == NA_OK) st = download_matrix(hc.s, c, (size_t)ldc, dc.as<double>(), lddc, m, n);
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) { *info = fail_info(st); return; }
    if (work) work[0] = 1.0;
}

// a (m x n) <- the first n columns of Q = H_1 ... H_k.
NAB_API void dorgqr_(const int* m_, const int* n_, const int* k_, double* a, const int* lda_, const double* tau, double* work,
                     const int* lwork_, int* info) {
    const int m = *m_, n = *n_, k = *k_, lda = *lda_, lwork = *lwork_;
    *info = 0;
    if (m < 0) { *info = -1; return; }
    if (n < 0 || n > m) { *info = -2; return; }
    if (k < 0 || k > n) { *info = -3; return; }
    if (lda < std::max(1, m)) { *info = -5; return; }
    if (lwork == -1) { if (work) work[0] = 1.0; return; }
    if (n == 0) return;
    HostCall hc;
    int st = hc.begin();
    if (st != NA_OK) { *info = fail_info(st); return; }
    Scratch d, dt, dq; size_t ldd;
    const size_t ldq = round_up((size_t)m, 2);
    st = upload_matrix(hc.s, d, ldd, a, (size_t)lda, m, std::max(k, 1));
    if (st == NA_OK) st = dq.alloc(ldq * n * sizeof(double), hc.s);
    if (st == NA_OK) st = set_identity(hc.s, dq.as<double>(), ldq, m, n);
    if (st == NA_OK && k > 0) {
        st = dt.alloc((size_t)k * sizeof(double), hc.s);
        if (st == NA_OK && cudaMemcpyAsync(dt.p, tau, (size_t)k * sizeof(double), cudaMemcpyHostToDevice, hc.s) != cudaSuccess) st = NA_ECUDA;
        if (st == NA_OK) st = apply_q_blocks(hc.s, m, k, d.as<double>(), ldd, nullptr, dq.as<double>(), ldq, n, false, n == k, dt.as<double>());
    }
    if (st == NA_OK) st = download_matrix(hc.s, a, (size_t)lda, dq.as<double>(), ldq, m, n);
    if (st != NA_OK || cudaStreamSynchronize(hc.s) != cudaSuccess) { *info = fail_info(st); return; }
    if (work) work[0] = 1.0;
}

// ---- triangular solve ------------------------------------------------------------------------------------------------
NAB_API void dtrtrs_(const char* uplo, const char* trans, const char* diag, const int* n_, const int* nrhs_, const double* a, const int* lda_,
                     double* b, const int* ldb_, int* info) {
    const int n = *n_, nrhs = *nrhs_, lda = *lda_, ldb = *ldb_;
    *info = 0;
    const bool lower = is(*uplo, 'L'), tr = is(*trans, 'T') || is(*trans, 'C'), unit = is(*diag, 'U');
    if (!(lower || is(*uplo, 'U'))) { *info = -1; return; }
    if (!(tr || is(*trans, 'N'))) { *info = -2; return; }
    if (!(unit || is(*diag, 'N'))) { *info = -3; return; }
    if (n < 0) { *info = -4; return; }
    if (nrhs < 0) { *info = -5; return; }
    if (lda < std::max(1, n)) { *info = -7; return; }
    if (ldb < std::max(1, n)) { *info = -9; return; }
    if (n == 0) return;
    if (!unit) for (int i = 0; i < n; ++i) if (a[i + (size_t)i * lda] == 0.0) { *info = i + 1; return; }   // exactly singular: b untouched
    if (nrhs == 0) return;
    const int st = na_tri_solve_f64(lower ? 1 : 0, tr ? 1 : 0, unit ? 1 : 0, (size_t)n, a, (size_t)lda, b, (size_t)ldb, (size_t)nrhs);
    if (st != NA_OK) *info = fail_info(st);
}

}  // extern "C"
