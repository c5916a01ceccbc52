// factor_qr.cu -- blocked Householder QR (compact WY) in nalgebra's storage, Q formation,
// Q^T application and QR solve, on the DGEMM tile engine.
//
// Reference: QR::new / q / q_tr_mul / solve_mut (/root/reference/src/linalg/qr.rs:55-129, 157-171,
// 204-256).  The reference applies one reflector at a time with dot + axpy per trailing column;
// here panels of 32 columns are factored by the cooperative GEQR2 kernel and applied as block
// reflectors I - V T V^T with three GEMMs; applying T (or T^T) is a triangular solve with
// S = T^-1 = triu(V^T V, 1) + diag(1/tau), so no LARFT recurrence is needed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    Scratch tau, vw, smat, wk, csign, ws_geqr2, ws_gram, ws_larfb;
    size_t ldv = 0, lds = 0, ldw = 0;
    int seq_state = 0, seq_larfb = 0;
    // storage conversion (and, for the host-pointer call, the download) of finished outer panels on a side stream
    StreamGuard sc;
    EventGuard ev_c;
    bool convert = false;             // false: the caller wants the classical form (tau_out)
    const QrHostSink* sink = nullptr;
    double* a = nullptr; size_t lda = 0, m = 0, n = 0, k = 0;
    double* diag = nullptr;
    size_t converted = 0;             // columns [0, converted) are done
    // Panel [j, j + jb) is final and nothing reads its classical form any more (the trailing updates use the clean
    // copy of V): convert it -- and every column left of it not yet converted -- behind the event recorded on `after`.
    int finish_columns(cudaStream_t after, size_t upto) {
        if (!convert || upto <= converted) return NA_OK;
        NAB_CUDA(cudaEventRecord(ev_c, after));
        NAB_CUDA(cudaStreamWaitEvent(sc, ev_c, 0));
        NAB_TRY(qr_convert_columns(sc, a, lda, m, k, converted, upto - converted, tau.as<double>(), csign.as<double>(), diag));
        if (sink)
            NAB_CUDA(cudaMemcpy2DAsync(sink->h + converted * sink->ldh, sink->ldh * 8, a + converted * lda, lda * 8, m * 8, upto - converted,
                                       cudaMemcpyDeviceToHost, sc));
        converted = upto;
        return NA_OK;
    }
    // the caller's stream continues after everything queued on the side stream
    int join(cudaStream_t s) {
        if (!convert) return NA_OK;
        NAB_CUDA(cudaEventRecord(ev_c, sc));
        NAB_CUDA(cudaStreamWaitEvent(s, ev_c, 0));
        return NA_OK;
    }
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
        NAB_TRY(ws_larfb.alloc(larfb_fused_workspace_bytes(), s));
        NAB_CUDA(cudaMemsetAsync(ws_larfb.p, 0, larfb_fused_workspace_bytes(), s));
        return NA_OK;
    }
};

// One outer panel: leaves of kQrLeaf columns (cooperative GEQR2), each applied to the rest of the panel as a
// 32-wide block reflector: one fused cooperative kernel (panel_qr_fused.cu) when the leaf is tall enough for its two
// grid-wide hand-offs to pay, the GEMM sequence otherwise.  vleaf / sleaf / wkleaf: leaf-level workspaces of the
// latter.  max_ctas: SMs the panel chain may occupy (0 = all).
static bool qr_fused_enabled();
// NAB_QR_LEAF=smem keeps the shared-memory GEQR2 leaf everywhere (A/B timing); default: the register-resident leaf
// (panel_qr_reg.cu) wherever its grid (512 rows per CTA) fits the SMs the panel may use.
static long g_qr_reg_leaf = [] { const char* e = getenv("NAB_QR_LEAF"); return (long)!(e && strcmp(e, "smem") == 0); }();
static long g_qr_fused = [] { const char* e = getenv("NAB_QR_FUSED"); return (long)(e ? atoi(e) != 0 : 1); }();
// na_set_tuning("qr_reg_leaf" / "qr_fused", 0 | 1): the tests factor the same matrix along every path
void qr_set_tuning(int which, long v) { (which == 0 ? g_qr_reg_leaf : g_qr_fused) = v; }
static bool qr_reg_leaf_enabled() { return g_qr_reg_leaf != 0; }
static bool qr_fused_enabled() { return g_qr_fused != 0; }
static bool qr_use_reg_leaf(size_t ml, int max_ctas) {
    if (!qr_reg_leaf_enabled() || ml < 4096) return false;
    const int g = geqr2_reg_grid(ml);
    return g > 0 && (max_ctas == 0 || g <= max_ctas);
}
static int qr_panel(cudaStream_t s, QrWork& w, size_t m, double* a, size_t lda, double* tau, size_t j, size_t jb,
                    double* vleaf, double* sleaf, double* wkleaf, int max_ctas = 0) {
    const size_t W = kQrLeaf;
    for (size_t l = 0; l < jb; l += W) {
        const size_t lw = std::min(W, jb - l), jl = j + l, ml = m - jl;
        double* apanel = a + jl + jl * lda;
        if (qr_use_reg_leaf(ml, max_ctas)) NAB_TRY(geqr2_panel_reg(s, apanel, lda, ml, lw, tau + jl, w.ws_geqr2.p, &w.seq_state));
        else NAB_TRY(geqr2_panel(s, apanel, lda, ml, lw, tau + jl, w.ws_geqr2.p, &w.seq_state));
        const size_t nc = (j + jb) - (jl + lw);          // rest of the outer panel
        if (nc > 0 && nc <= 224 && ml >= 2048 && qr_fused_enabled()) {
            NAB_TRY(larfb_leaf_fused(s, apanel, lda, ml, lw, nc, tau + jl, w.ws_larfb.p, &w.seq_larfb, max_ctas));
        } else if (nc > 0) {
            NAB_TRY(extract_v_gram(s, vleaf, w.ldv, apanel, lda, ml, lw, tau + jl, sleaf, w.lds, w.ws_gram.p));
            NAB_TRY(apply_block_reflector(s, ml, lw, vleaf, w.ldv, sleaf, w.lds, true, apanel + lw * lda, lda, nc, wkleaf, w.ldw));
        }
    }
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Tall matrices with many panels: right-looking outer blocks with one-step look-ahead on two streams, like
// the LU driver.  la(j) applies block reflector j to the next panel's columns (whole GPU, critical chain);
// bulk(j) applies it to the columns right of that: part A beside panel(j + nb) on the SMs the cooperative
// GEQR2 grid leaves free, part B on the whole GPU once the panel is done.  The outer V / S are double-buffered
// (bulk B(j) may still read them while panel(j + nb) builds its own leaves' V / S in separate buffers).
// ------------------------------------------------------------------------------------------------
static bool qr_lookahead_enabled() {
    static bool v = [] { const char* e = getenv("NAB_QR_LOOKAHEAD"); return e ? atoi(e) != 0 : true; }();
    return v;
}

static int qr_lookahead(cudaStream_t sp, QrWork& w, size_t m, size_t n, double* a, size_t lda, double* tau, size_t k) {
    cudaStream_t su = nullptr;
    cudaEvent_t ev_p = nullptr, ev_u = nullptr, ev_d = nullptr;
    NAB_CUDA(cudaStreamCreateWithFlags(&su, cudaStreamNonBlocking));
    NAB_CUDA(cudaEventCreateWithFlags(&ev_p, cudaEventDisableTiming));
    NAB_CUDA(cudaEventCreateWithFlags(&ev_u, cudaEventDisableTiming));
    NAB_CUDA(cudaEventCreateWithFlags(&ev_d, cudaEventDisableTiming));
    Scratch vwo[2], smo[2], wkla, wkbulk;
    int st = NA_OK;
    for (int i = 0; i < 2 && st == NA_OK; ++i) {
        st = vwo[i].alloc(w.ldv * QR_NB * sizeof(double), sp);
        if (st == NA_OK) st = smo[i].alloc(w.lds * QR_NB * sizeof(double), sp);
    }
    if (st == NA_OK) st = wkla.alloc(w.ldw * QR_NB * sizeof(double), sp);
    if (st == NA_OK) st = wkbulk.alloc(w.ldw * std::max<size_t>(n, 1) * sizeof(double), sp);
    const int sms = ctx().sm_count;
    Timeline tr("NAB_QR_TRACE", "qr_trace");
    tr.start(sp);
    auto outer = [&](cudaStream_t s, size_t j, size_t jb, int par, size_t c0, size_t nc, double* wk) {
        // columns [c0, c0 + nc) <- (I - V T^T V^T) columns, V = outer panel j (rows j..)
        return apply_block_reflector(s, m - j, jb, vwo[par].as<double>(), w.ldv, smo[par].as<double>(), w.lds, true,
                                     a + j + c0 * lda, lda, nc, wk, w.ldw);
    };
    auto prep = [&](size_t j, size_t jb, int par) {
        NAB_TRY(extract_v(sp, vwo[par].as<double>(), w.ldv, a + j + j * lda, lda, m - j, jb, tau + j, 0));
        return build_s_from_v(sp, m - j, jb, vwo[par].as<double>(), w.ldv, tau + j, smo[par].as<double>(), w.lds);
    };
    bool bulk_pending = false;
    int par = 0;
    cudaEvent_t t0 = tr.mark(sp);
    if (st == NA_OK) st = qr_panel(sp, w, m, a, lda, tau, 0, std::min(QR_NB, k), w.vw.as<double>(), w.smat.as<double>(), w.wk.as<double>());
    tr.add("panel", 0, t0, tr.mark(sp));
    for (size_t j = 0; st == NA_OK; j += QR_NB) {
        const size_t jb = std::min(QR_NB, k - j), jn = j + jb;
        const size_t jbn = jn < k ? std::min(QR_NB, k - jn) : 0;
        if (jn >= n) break;                                   // nothing right of this panel
        st = prep(j, jb, par);
        if (st == NA_OK) st = w.finish_columns(sp, jn);       // panel j is final; the updates read the clean copy of V
        if (st != NA_OK) break;
        if (bulk_pending) cudaStreamWaitEvent(sp, ev_u, 0);
        // la(j): the next panel's columns (or, after the last panel, nothing)
        cudaEvent_t t_la = tr.mark(sp);
        if (jbn) st = outer(sp, j, jb, par, jn, jbn, wkla.as<double>());
        if (st != NA_OK) break;
        tr.add("la", j, t_la, tr.mark(sp));
        cudaEventRecord(ev_p, sp);
        const size_t x0 = jn + jbn, nx = n - x0, mj = m - j;
        // SMs of the panel chain = the cooperative GEQR2 grid of the next panel's first (tallest) leaf
        // (the register-resident leaf takes one CTA per 512 rows: nearly the whole GPU for a 65536-row panel, which then
        // runs much faster than beside a large bulk share)
        const int rp = !jbn ? 0
                     : qr_use_reg_leaf(m - jn, 0) ? std::min(sms - 8, geqr2_reg_grid(m - jn) + 2)
                                                  : std::min(sms - 16, geqr2_grid(m - jn, std::min<size_t>(kQrLeaf, jbn)) + 2);
        size_t wa = nx;
        if (jbn && nx) {
            // measured per 256 columns: ~3.5 ms with the shared-memory leaf beside the bulk update, ~2.3 ms with the register leaf
            const double t_panel = (double)jbn / 32.0 * (qr_use_reg_leaf(m - jn, 0) ? 0.28e-3 : 0.44e-3);
            const double target = t_panel * (sms - rp) * 0.18e12;
            wa = round_up((size_t)(target / (4.0 * (double)mj * (double)jb)) + 1, 128);
            if (wa + 256 >= nx) wa = nx;
        }
        if (nx) {
            cudaStreamWaitEvent(su, ev_p, 0);
            cudaEvent_t t_a = tr.mark(su);
            set_gemm_sm_limit(jbn ? sms - rp : 0);
            st = outer(su, j, jb, par, x0, wa, wkbulk.as<double>());
            set_gemm_sm_limit(0);
            if (st != NA_OK) break;
            tr.add("bulkA", j, t_a, tr.mark(su));
        }
        if (jbn == 0) { if (nx) { cudaEventRecord(ev_u, su); bulk_pending = true; } break; }
        cudaEvent_t t_p = tr.mark(sp);
        set_gemm_sm_limit(rp);
        st = qr_panel(sp, w, m, a, lda, tau, jn, jbn, w.vw.as<double>(), w.smat.as<double>(), w.wk.as<double>(), rp);
        set_gemm_sm_limit(0);
        if (st != NA_OK) break;
        tr.add("panel", jn, t_p, tr.mark(sp));
        if (nx) {
            if (wa < nx) {
                cudaEventRecord(ev_d, sp);
                cudaStreamWaitEvent(su, ev_d, 0);
                cudaEvent_t t_b = tr.mark(su);
                st = outer(su, j, jb, par, x0 + wa, nx - wa, wkbulk.as<double>() + wa * w.ldw);
                if (st != NA_OK) break;
                tr.add("bulkB", j, t_b, tr.mark(su));
            }
            cudaEventRecord(ev_u, su);
            bulk_pending = true;
        }
        par ^= 1;
    }
    if (bulk_pending) cudaStreamWaitEvent(sp, ev_u, 0);
    cudaStreamSynchronize(su);
    tr.dump();
    cudaEventDestroy(ev_p); cudaEventDestroy(ev_u); cudaEventDestroy(ev_d); cudaStreamDestroy(su);
    return st;
}

// diag: DEVICE pointer, min(m,n) entries (nalgebra storage).  tau_out != nullptr instead: stop before the storage
// conversion -- `a` then holds the classical / LAPACK geqrf form (R on and above the diagonal, the reflector vectors v
// with implicit unit head below it) and tau_out (DEVICE, min(m,n)) the scalar factors.
int qr_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diag, double* tau_out, const QrHostSink* sink) {
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    if (lda < m) { set_error("qr: lda < m"); return NA_EINVAL; }
    if (m > 0x7fffff00ull || n > 0x7fffff00ull) { set_error("qr: dimension exceeds 2^31"); return NA_EINVAL; }
    QrWork w;
    NAB_TRY(w.init(s, m, n, k));
    double* tau = w.tau.as<double>();
    double* vw = w.vw.as<double>();
    if (!tau_out) {
        // nalgebra's storage is produced panel by panel on a side stream while the factorization goes on
        NAB_TRY(w.sc.create()); NAB_TRY(w.ev_c.create());
        w.convert = true; w.sink = sink;
        w.a = a; w.lda = lda; w.m = m; w.n = n; w.k = k; w.diag = diag;
    }
    auto finish = [&]() -> int {
        if (tau_out) { NAB_CUDA(cudaMemcpyAsync(tau_out, tau, k * sizeof(double), cudaMemcpyDeviceToDevice, s)); return NA_OK; }
        NAB_TRY(w.finish_columns(s, n));                      // whatever is left: the last panel, columns right of min(m, n)
        return w.join(s);
    };
    // Look-ahead pays when the panel leaves SMs to the bulk update (shared-memory leaf: ~98 CTAs at 65536 rows).  The
    // register-resident leaf takes nearly the whole GPU and is 2x faster, so the plain loop is the better schedule with it
    // (65536 x 4096: 102 ms against 112 ms with look-ahead and 106 ms with the shared-memory leaf + look-ahead).
    if (qr_lookahead_enabled() && k >= 4 * QR_NB && m >= 8192 && n <= m && !qr_use_reg_leaf(m, 0)) {
        NAB_TRY(qr_lookahead(s, w, m, n, a, lda, tau, k));
        return finish();
    }
    for (size_t j = 0; j < k; j += QR_NB) {
        const size_t jb = std::min(QR_NB, k - j), mj = m - j;
        NAB_TRY(qr_panel(s, w, m, a, lda, tau, j, jb, vw, w.smat.as<double>(), w.wk.as<double>()));
        const size_t nt = n - (j + jb);                       // trailing columns
        if (nt > 0) {
            double* apanel = a + j + j * lda;
            NAB_TRY(extract_v(s, vw, w.ldv, apanel, lda, mj, jb, tau + j, 0));
            NAB_TRY(build_s_from_v(s, mj, jb, vw, w.ldv, tau + j, w.smat.as<double>(), w.lds));
            NAB_TRY(w.finish_columns(s, j + jb));             // the updates below read the clean copy of V, not the panel
            NAB_TRY(apply_block_reflector(s, mj, jb, vw, w.ldv, w.smat.as<double>(), w.lds, true,
                                          apanel + jb * lda, lda, nt, w.wk.as<double>(), w.ldw));
        }
    }
    return finish();
}

// B (m x nb, ldb) <- Q_L^T B (forward = true) or Q_L B restricted to the trailing block structure
// (forward = false), where Q_L = H_0 H_1 ... H_{k-1}, H_i = I - 2 u_i u_i^T from nalgebra's storage.
// For forward == false only columns [col_lo(j), nb) of B are touched at block j when `triangular_q`
// (forming Q from [I; 0]: columns left of the block are still unit vectors untouched by it).
// lapack_tau != nullptr: `qr` is in LAPACK geqrf storage with these scalar factors (DEVICE) instead of nalgebra's.
int apply_q_blocks(cudaStream_t s, size_t m, size_t k, const double* qr, size_t lda, const double* diag,
                   double* b, size_t ldb, size_t nb, bool forward, bool triangular_q, const double* lapack_tau) {
    if (k == 0 || nb == 0) return NA_OK;
    QrWork w;
    NAB_TRY(w.init(s, m, nb, k));
    double* tau = w.tau.as<double>();
    if (lapack_tau) NAB_CUDA(cudaMemcpyAsync(tau, lapack_tau, k * sizeof(double), cudaMemcpyDeviceToDevice, s));
    else NAB_TRY(tau_from_diag(s, tau, diag, k));
    const size_t nblk = ceil_div(k, QR_NB);
    for (size_t bi = 0; bi < nblk; ++bi) {
        const size_t blk = forward ? bi : nblk - 1 - bi;
        const size_t j = blk * QR_NB, jb = std::min(QR_NB, k - j), mj = m - j;
        NAB_TRY(extract_v(s, w.vw.as<double>(), w.ldv, qr + j + j * lda, lda, mj, jb, tau + j, lapack_tau ? 0 : 1));
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
    NAB_TRY(apply_q_blocks(s, m, k, qr, lda, diag, q, ldq, k, false, true, nullptr));
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
    NAB_TRY(apply_q_blocks(s, m, k, qr, lda, diag, b, ldb, nrhs, true, false, nullptr));
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
    return qr_device(static_cast<cudaStream_t>(stream), m, n, a, lda, diag, nullptr);
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
    // finished column blocks go back over PCIe behind the factorization (overlap needs pinned host memory)
    QrHostSink sink{a, lda};
    const int st = qr_device(s, m, n, d.as<double>(), ldd, dd.as<double>(), nullptr, &sink);
    if (st < 0) { cudaStreamSynchronize(s); return st; }
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
