// factor_qr.cu -- blocked Householder QR (compact WY) in nalgebra's storage, Q formation,
// Q^T application and QR solve, on the DGEMM tile engine.
//
// Reference: QR::new / q / q_tr_mul / solve_mut (/root/reference/src/linalg/qr.rs:55-129, 157-171,
// 204-256).  The reference applies one reflector at a time with dot + axpy per trailing column;
// here panels of 32 columns are factored by the cooperative GEQR2 kernel and applied as block
// reflectors I - V T V^T with three GEMMs; applying T (or T^T) is a triangular solve with
// S = T^-1 = triu(V^T V, 1) + diag(1/tau), so no LARFT recurrence is needed.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace nab {

constexpr size_t QR_NB = 256;            // outer block (K of the big trailing GEMMs)

// C (mc x nc, ldc) <- (I - V T^(T) V^T) C   with S = T^-1 given (w x w, lds); V is mc x w (ldv).
// transpose_t: true applies T^T (this is Q^T = H_k..H_1 for a forward block), false applies T.
static int apply_block_reflector(cudaStream_t s, size_t mc, size_t w, const double* v, size_t ldv, const double* smat, size_t lds,
                                 bool transpose_t, double* c, size_t ldc, size_t nc, double* wk, size_t ldw) {
    if (mc == 0 || nc == 0 || w == 0) return NA_OK;
    // W = V^T C   (w x nc), K = mc: split-K inside the GEMM engine
    NAB_TRY(dgemm_device(s, false, w, mc, nc, 1.0, v, (ptrdiff_t)ldv, 1, c, 1, (ptrdiff_t)ldc, 0.0, wk, 1, (ptrdiff_t)ldw));
    // W <- T^T W = S^-T W  (solve S^T X = W, S^T lower)   or   W <- T W = S^-1 W (solve S X = W, S upper)
    if (transpose_t) NAB_TRY(trsm_left(s, true, false, w, smat, (ptrdiff_t)lds, 1, nullptr, nullptr, wk, 1, (ptrdiff_t)ldw, nc));
    else NAB_TRY(trsm_left(s, false, false, w, smat, 1, (ptrdiff_t)lds, nullptr, nullptr, wk, 1, (ptrdiff_t)ldw, nc));
    // C -= V W
    return dgemm_device(s, false, mc, w, nc, -1.0, v, 1, (ptrdiff_t)ldv, wk, 1, (ptrdiff_t)ldw, 1.0, c, 1, (ptrdiff_t)ldc);
}

// S = triu(V^T V, 1) + diag(1/tau)
static int build_s_from_v(cudaStream_t s, size_t mc, size_t w, const double* v, size_t ldv, const double* tau, double* smat, size_t lds) {
    NAB_TRY(dgemm_device(s, false, w, mc, w, 1.0, v, (ptrdiff_t)ldv, 1, v, 1, (ptrdiff_t)ldv, 0.0, smat, 1, (ptrdiff_t)lds));
    return build_s(s, smat, lds, w, tau);
}

struct QrWork {
    Scratch tau, vw, smat, wk, csign, ws_geqr2, ws_gram;
    size_t ldv = 0, lds = 0, ldw = 0;
    int seq_state = 0;
    int init(cudaStream_t s, size_t m, size_t ncols_max, size_t k) {
        ldv = round_up(m, 2); lds = QR_NB; ldw = QR_NB;
        NAB_TRY(tau.alloc(std::max<size_t>(k, 1) * sizeof(double), s));
        NAB_TRY(vw.alloc(ldv * QR_NB * sizeof(double), s));
        NAB_TRY(smat.alloc(lds * QR_NB * sizeof(double), s));
        NAB_TRY(wk.alloc(ldw * std::max<size_t>(ncols_max, 1) * sizeof(double), s));
        NAB_TRY(csign.alloc((k + 2) * sizeof(double), s));
        NAB_TRY(ws_geqr2.alloc(geqr2_workspace_bytes(), s));
        NAB_CUDA(cudaMemsetAsync(ws_geqr2.p, 0, geqr2_workspace_bytes(), s));
        NAB_TRY(ws_gram.alloc(extract_v_gram_workspace_bytes(), s));
        NAB_CUDA(cudaMemsetAsync(ws_gram.p, 0, extract_v_gram_workspace_bytes(), s));
        return NA_OK;
    }
};

// diag: DEVICE pointer, min(m,n) entries.
int qr_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diag) {
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    if (lda < m) { set_error("qr: lda < m"); return NA_EINVAL; }
    if (m > 0x7fffff00ull || n > 0x7fffff00ull) { set_error("qr: dimension exceeds 2^31"); return NA_EINVAL; }
    QrWork w;
    NAB_TRY(w.init(s, m, n, k));
    double* tau = w.tau.as<double>();
    double* vw = w.vw.as<double>();
    const size_t W = kQrLeaf;
    for (size_t j = 0; j < k; j += QR_NB) {
        const size_t jb = std::min(QR_NB, k - j), mj = m - j;
        for (size_t l = 0; l < jb; l += W) {
            const size_t lw = std::min(W, jb - l), jl = j + l, ml = m - jl;
            double* apanel = a + jl + jl * lda;
            NAB_TRY(geqr2_panel(s, apanel, lda, ml, lw, tau + jl, w.ws_geqr2.p, &w.seq_state));
            const size_t nc = (j + jb) - (jl + lw);          // rest of the outer panel
            if (nc > 0) {
                NAB_TRY(extract_v_gram(s, vw, w.ldv, apanel, lda, ml, lw, tau + jl, w.smat.as<double>(), w.lds, w.ws_gram.p));
                NAB_TRY(apply_block_reflector(s, ml, lw, vw, w.ldv, w.smat.as<double>(), w.lds, true,
                                              apanel + lw * lda, lda, nc, w.wk.as<double>(), w.ldw));
            }
        }
        const size_t nt = n - (j + jb);                       // trailing columns
        if (nt > 0) {
            double* apanel = a + j + j * lda;
            NAB_TRY(extract_v(s, vw, w.ldv, apanel, lda, mj, jb, tau + j, 0));
            NAB_TRY(build_s_from_v(s, mj, jb, vw, w.ldv, tau + j, w.smat.as<double>(), w.lds));
            NAB_TRY(apply_block_reflector(s, mj, jb, vw, w.ldv, w.smat.as<double>(), w.lds, true,
                                          apanel + jb * lda, lda, nt, w.wk.as<double>(), w.ldw));
        }
    }
    return qr_convert_to_nalgebra(s, a, lda, m, n, tau, w.csign.as<double>(), diag);
}

// B (m x nb, ldb) <- Q_L^T B (forward = true) or Q_L B restricted to the trailing block structure
// (forward = false), where Q_L = H_0 H_1 ... H_{k-1}, H_i = I - 2 u_i u_i^T from nalgebra's storage.
// For forward == false only columns [col_lo(j), nb) of B are touched at block j when `triangular_q`
// (forming Q from [I; 0]: columns left of the block are still unit vectors untouched by it).
static int apply_q_blocks(cudaStream_t s, size_t m, size_t k, const double* qr, size_t lda, const double* diag,
                          double* b, size_t ldb, size_t nb, bool forward, bool triangular_q) {
    if (k == 0 || nb == 0) return NA_OK;
    QrWork w;
    NAB_TRY(w.init(s, m, nb, k));
    double* tau = w.tau.as<double>();
    NAB_TRY(tau_from_diag(s, tau, diag, k));
    const size_t nblk = ceil_div(k, QR_NB);
    for (size_t bi = 0; bi < nblk; ++bi) {
        const size_t blk = forward ? bi : nblk - 1 - bi;
        const size_t j = blk * QR_NB, jb = std::min(QR_NB, k - j), mj = m - j;
        NAB_TRY(extract_v(s, w.vw.as<double>(), w.ldv, qr + j + j * lda, lda, mj, jb, tau + j, 1));
        NAB_TRY(build_s_from_v(s, mj, jb, w.vw.as<double>(), w.ldv, tau + j, w.smat.as<double>(), w.lds));
        const size_t c0 = triangular_q ? j : 0;
        NAB_TRY(apply_block_reflector(s, mj, jb, w.vw.as<double>(), w.ldv, w.smat.as<double>(), w.lds, forward,
                                      b + j + c0 * ldb, ldb, nb - c0, w.wk.as<double>(), w.ldw));
    }
    return NA_OK;
}

// QR::q: q (m x k, ldq) on device.
int qr_q_device(cudaStream_t s, size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq) {
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    NAB_TRY(set_identity(s, q, ldq, m, k));
    NAB_TRY(apply_q_blocks(s, m, k, qr, lda, diag, q, ldq, k, false, true));
    Scratch csign;
    NAB_TRY(csign.alloc((k + 2) * sizeof(double), s));
    NAB_TRY(qr_signs_from_diag(s, diag, k, csign.as<double>()));
    return scale_signs(s, q, ldq, m, k, csign.as<double>(), k, true);
}

// QR::q_tr_mul: b (m x nrhs) <- Q^T b.
int qr_q_tr_mul_device(cudaStream_t s, size_t m, size_t n, const double* qr, size_t lda, const double* diag,
                       double* b, size_t ldb, size_t nrhs) {
    const size_t k = std::min(m, n);
    if (k == 0 || nrhs == 0) return NA_OK;
    NAB_TRY(apply_q_blocks(s, m, k, qr, lda, diag, b, ldb, nrhs, true, false));
    Scratch csign;
    NAB_TRY(csign.alloc((k + 2) * sizeof(double), s));
    NAB_TRY(qr_signs_from_diag(s, diag, k, csign.as<double>()));
    return scale_signs(s, b, ldb, m, nrhs, csign.as<double>(), k, false);
}

// QR::solve_mut (square): q_tr_mul, then back substitution with R (diagonal = |diag|).
int qr_solve_device(cudaStream_t s, size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs) {
    if (n == 0 || nrhs == 0) return NA_OK;
    NAB_TRY(qr_q_tr_mul_device(s, n, n, qr, lda, diag, b, ldb, nrhs));
    Scratch flag;
    NAB_TRY(flag.alloc(sizeof(int), s));
    NAB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
    NAB_TRY(zero_diag_check(s, nullptr, 0, diag, n, flag.as<int>()));
    int h = 0;
    NAB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    if (h) return NA_SINGULAR;                                   // qr.rs:242-244
    return trsm_left(s, false, false, n, qr, 1, (ptrdiff_t)lda, diag, nullptr, b, 1, (ptrdiff_t)ldb, nrhs);
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_qr_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diag, void* stream) {
    NAB_TRY(ensure_init());
    return qr_device(static_cast<cudaStream_t>(stream), m, n, a, lda, diag);
}

int na_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag) {
    NAB_TRY(ensure_init());
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    if (!a || !diag || lda < m) { set_error("qr: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d, dd; size_t ldd;
    NAB_TRY(upload_matrix(s, d, ldd, a, lda, m, n));
    NAB_TRY(dd.alloc(k * sizeof(double), s));
    NAB_TRY(qr_device(s, m, n, d.as<double>(), ldd, dd.as<double>()));
    NAB_TRY(download_matrix(s, a, lda, d.as<double>(), ldd, m, n));
    NAB_CUDA(cudaMemcpyAsync(diag, dd.p, k * sizeof(double), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

int na_qr_q_f64_dev(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq, void* stream) {
    NAB_TRY(ensure_init());
    return qr_q_device(static_cast<cudaStream_t>(stream), m, n, qr, lda, diag, q, ldq);
}

int na_qr_q_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* q, size_t ldq) {
    NAB_TRY(ensure_init());
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    if (!qr || !diag || !q || lda < m || ldq < m) { set_error("qr_q: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d, dd, dq; size_t ldd;
    NAB_TRY(upload_matrix(s, d, ldd, qr, lda, m, n));
    NAB_TRY(dd.alloc(k * sizeof(double), s));
    NAB_CUDA(cudaMemcpyAsync(dd.p, diag, k * sizeof(double), cudaMemcpyHostToDevice, s));
    const size_t lddq = round_up(m, 2);
    NAB_TRY(dq.alloc(lddq * k * sizeof(double), s));
    NAB_TRY(qr_q_device(s, m, n, d.as<double>(), ldd, dd.as<double>(), dq.as<double>(), lddq));
    NAB_TRY(download_matrix(s, q, ldq, dq.as<double>(), lddq, m, k));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

int na_qr_q_tr_mul_f64_dev(size_t m, size_t n, const double* qr, size_t lda, const double* diag,
                           double* b, size_t ldb, size_t nrhs, void* stream) {
    NAB_TRY(ensure_init());
    return qr_q_tr_mul_device(static_cast<cudaStream_t>(stream), m, n, qr, lda, diag, b, ldb, nrhs);
}

static int qr_host_apply(bool solve, size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs) {
    const size_t k = std::min(m, n);
    if (k == 0 || nrhs == 0) return NA_OK;
    if (!qr || !diag || !b || lda < m || ldb < m) { set_error("qr apply: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d, dd, db; size_t ldd, lddb;
    NAB_TRY(upload_matrix(s, d, ldd, qr, lda, m, n));
    NAB_TRY(dd.alloc(k * sizeof(double), s));
    NAB_CUDA(cudaMemcpyAsync(dd.p, diag, k * sizeof(double), cudaMemcpyHostToDevice, s));
    NAB_TRY(upload_matrix(s, db, lddb, b, ldb, m, nrhs));
    int st = solve ? qr_solve_device(s, m, d.as<double>(), ldd, dd.as<double>(), db.as<double>(), lddb, nrhs)
                   : qr_q_tr_mul_device(s, m, n, d.as<double>(), ldd, dd.as<double>(), db.as<double>(), lddb, nrhs);
    if (st < 0) return st;
    NAB_TRY(download_matrix(s, b, ldb, db.as<double>(), lddb, m, nrhs));
    NAB_CUDA(cudaStreamSynchronize(s));
    return st;
}

int na_qr_q_tr_mul_f64(size_t m, size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs) {
    NAB_TRY(ensure_init());
    return qr_host_apply(false, m, n, qr, lda, diag, b, ldb, nrhs);
}

int na_qr_solve_f64(size_t n, const double* qr, size_t lda, const double* diag, double* b, size_t ldb, size_t nrhs) {
    NAB_TRY(ensure_init());
    return qr_host_apply(true, n, n, qr, lda, diag, b, ldb, nrhs);
}

}  // extern "C"
