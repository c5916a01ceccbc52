// factor_chol.cu -- blocked triangular solves and the blocked Cholesky on the DGEMM tile engine,
// plus their extern "C" entry points.
//
// Reference semantics: Cholesky::new / new_with_substitute / solve_mut
// (/root/reference/src/linalg/cholesky.rs:196-272, 122-129) and the triangular solves of
// /root/reference/src/linalg/solve.rs.  The reference runs unblocked Level-1 loops; here the
// O(n^3) work is restructured as recursive panel + TRSM + SYRK so that it runs on the DMMA GEMM
// engine.  Parity is therefore on results (residuals, failure column), not on operation order.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace nab {

constexpr size_t IBs = kInvBlock;

// ------------------------------------------------------------------------------------------------
// TRSM (left side): M X = B, recursive; base case X_b = inv(M_bb) * B_b through the GEMM engine.
// ------------------------------------------------------------------------------------------------
struct TrsmCtx {
    cudaStream_t s;
    bool eff_lower;
    const double* m; ptrdiff_t rsm, csm;
    const double* inv;            // [nblk][128*128]
    double* b; ptrdiff_t rsb, csb; size_t nrhs;
    double* tmp; ptrdiff_t rst, cst;   // 128 x nrhs scratch with B's orientation
};

static int trsm_rec(const TrsmCtx& c, size_t r0, size_t len) {
    if (len == 0) return NA_OK;
    if (len <= IBs) {
        const double* inv = c.inv + (r0 / IBs) * IBs * IBs;
        double* brow = c.b + (ptrdiff_t)r0 * c.rsb;
        NAB_TRY(dgemm_device(c.s, false, len, len, c.nrhs, 1.0, inv, 1, (ptrdiff_t)IBs, brow, c.rsb, c.csb, 0.0, c.tmp, c.rst, c.cst));
        return copy_strided(c.s, brow, c.rsb, c.csb, c.tmp, c.rst, c.cst, len, c.nrhs);
    }
    const size_t h = round_up(len / 2, IBs);
    const size_t r1 = r0 + h, len1 = len - h;
    if (c.eff_lower) {
        NAB_TRY(trsm_rec(c, r0, h));
        // B[r1:] -= M[r1:, r0:r1] * B[r0:r1]
        NAB_TRY(dgemm_device(c.s, false, len1, h, c.nrhs, -1.0, c.m + (ptrdiff_t)r1 * c.rsm + (ptrdiff_t)r0 * c.csm, c.rsm, c.csm,
                             c.b + (ptrdiff_t)r0 * c.rsb, c.rsb, c.csb, 1.0, c.b + (ptrdiff_t)r1 * c.rsb, c.rsb, c.csb));
        return trsm_rec(c, r1, len1);
    } else {
        NAB_TRY(trsm_rec(c, r1, len1));
        // B[r0:r1] -= M[r0:r1, r1:] * B[r1:]
        NAB_TRY(dgemm_device(c.s, false, h, len1, c.nrhs, -1.0, c.m + (ptrdiff_t)r0 * c.rsm + (ptrdiff_t)r1 * c.csm, c.rsm, c.csm,
                             c.b + (ptrdiff_t)r1 * c.rsb, c.rsb, c.csb, 1.0, c.b + (ptrdiff_t)r0 * c.rsb, c.rsb, c.csb));
        return trsm_rec(c, r0, h);
    }
}

int trsm_left(cudaStream_t s, bool eff_lower, bool unit, size_t n, const double* m, ptrdiff_t rsm, ptrdiff_t csm,
              const double* diag_abs, const double* inv_blocks, double* b, ptrdiff_t rsb, ptrdiff_t csb, size_t nrhs) {
    if (n == 0 || nrhs == 0) return NA_OK;
    Scratch inv, tmp;
    if (!inv_blocks) {
        NAB_TRY(inv.alloc(ceil_div(n, IBs) * IBs * IBs * sizeof(double), s));
        NAB_TRY(trtri_blocks(s, m, rsm, csm, n, eff_lower, unit, diag_abs, inv.as<double>()));
        inv_blocks = inv.as<double>();
    }
    TrsmCtx c{s, eff_lower, m, rsm, csm, inv_blocks, b, rsb, csb, nrhs, nullptr, 0, 0};
    const bool b_row_fast = (rsb == 1);
    const size_t ldt = b_row_fast ? IBs : round_up(nrhs, 2);
    NAB_TRY(tmp.alloc((b_row_fast ? ldt * nrhs : ldt * IBs) * sizeof(double), s));
    c.tmp = tmp.as<double>();
    c.rst = b_row_fast ? 1 : (ptrdiff_t)ldt;
    c.cst = b_row_fast ? (ptrdiff_t)ldt : 1;
    return trsm_rec(c, 0, n);
}

// ------------------------------------------------------------------------------------------------
// Cholesky: recursive, leaves = POTF2 on 128-blocks (+ TRTRI of the leaf for the TRSMs above it).
// ------------------------------------------------------------------------------------------------
struct CholCtx {
    cudaStream_t s;
    double* a; size_t lda;
    int use_sub; double sub;
    double* inv;                      // [n/128][128*128] inverses of L's diagonal blocks
    unsigned long long* fail;         // device: first failing column (init ~0)
    size_t col_offset = 0;            // added to the reported failing column (block-cyclic drivers: global column)
    const struct HostSink* sink = nullptr;   // host-pointer call: finished column blocks stream out behind the panels
};

// The host-pointer entry point hands this to the look-ahead driver: as soon as panel j is final (nothing touches the
// columns of L left of the trailing matrix again) its lower trapezoid goes back over PCIe on a copy stream while the
// trailing update and the next panels run.  Only the lower triangle crosses the bus, in both directions.
struct HostSink {
    double* h; size_t ldh;
    cudaStream_t sc;                  // copy stream
    cudaEvent_t ev;                   // reused: record on the compute stream, wait on the copy stream
};
static int sink_panel(const CholCtx& c, cudaStream_t sp, size_t n, size_t j, size_t jb) {
    if (!c.sink) return NA_OK;
    NAB_CUDA(cudaEventRecord(c.sink->ev, sp));
    NAB_CUDA(cudaStreamWaitEvent(c.sink->sc, c.sink->ev, 0));
    NAB_CUDA(cudaMemcpy2DAsync(c.sink->h + j + j * c.sink->ldh, c.sink->ldh * 8, c.a + j + j * c.lda, c.lda * 8, (n - j) * 8, jb,
                               cudaMemcpyDeviceToHost, c.sink->sc));
    return NA_OK;
}

static int chol_rec(const CholCtx& c, size_t j0, size_t n) {
    if (n == 0) return NA_OK;
    double* ajj = c.a + j0 + j0 * c.lda;
    if (n <= IBs) {
        return potf2(c.s, ajj, c.lda, (int)n, c.use_sub, c.sub, j0 + c.col_offset, c.fail, c.inv + (j0 / IBs) * IBs * IBs);
    }
    const size_t n1 = round_up(n / 2, IBs), n2 = n - n1;
    NAB_TRY(chol_rec(c, j0, n1));
    double* a21 = ajj + n1;
    // A21 <- A21 * L11^-T   <=>   L11 * X^T = A21^T  (left solve on the transposed view of A21)
    NAB_TRY(trsm_left(c.s, true, false, n1, ajj, 1, (ptrdiff_t)c.lda, nullptr, c.inv + (j0 / IBs) * IBs * IBs,
                      a21, (ptrdiff_t)c.lda, 1, n2));
    // A22 <- A22 - A21 * A21^T, lower triangle only
    double* a22 = ajj + n1 + n1 * c.lda;
    NAB_TRY(dgemm_device(c.s, true, n2, n1, n2, -1.0, a21, 1, (ptrdiff_t)c.lda, a21, (ptrdiff_t)c.lda, 1, 1.0, a22, 1, (ptrdiff_t)c.lda));
    return chol_rec(c, j0 + n1, n2);
}

// ------------------------------------------------------------------------------------------------
// Large n: right-looking blocked Cholesky with one-step look-ahead on two streams.
//
//   panel(j)   : columns [j, j+nb): inner right-looking loop over 128-blocks -- POTF2+TRTRI leaf,
//                TRSM as one GEMM with the leaf's inverse, lower-trapezoid GEMM on the rest of the panel.
//   la(j)      : update of the NEXT panel's columns with panel j (trapezoid GEMM, K = nb), whole GPU.
//   bulk(j)    : update of everything right of the next panel with panel j (lower-only GEMM, K = nb).
//
// bulk(j) runs on a second stream on (SMs - Rp) CTAs while panel(j+nb) runs on the caller's stream on
// at most Rp CTAs, so the latency-bound panel is hidden behind the compute-bound update; the GEMM
// kernels are persistent (one CTA per SM), which is why the split is expressed as grid limits.
// ------------------------------------------------------------------------------------------------
static size_t chol_nb() {
    static size_t v = [] { const char* e = getenv("NAB_CHOL_NB"); size_t x = e ? (size_t)atoi(e) : 512; return x >= 128 ? x / 128 * 128 : 512; }();
    return v;
}
static int chol_rp_override() {
    static int v = [] { const char* e = getenv("NAB_CHOL_RP"); return e ? atoi(e) : 0; }();
    return v;
}
// Off by default: at N = 16384 the panel chain is as long as the bulk update in most steps, so part B would
// mostly wait for the panel (measured 60.9 ms with the split vs 56.9 ms without).
static bool chol_split() {
    static bool v = [] { const char* e = getenv("NAB_CHOL_SPLIT"); return e ? atoi(e) != 0 : false; }();
    return v;
}
static int chol_rp_min() {
    static int v = [] { const char* e = getenv("NAB_CHOL_RP_MIN"); return e ? std::max(4, atoi(e)) : 12; }();     // measured at 16384: 16 -> 54.7 ms, 12 -> 54.05, 8 -> 54.2
    return v;
}
#define CHOL_NB (chol_nb())

static int chol_panel(const CholCtx& c, cudaStream_t sp, size_t n, size_t j, size_t jb, int sm_limit, double* tmp, size_t ldt) {
    for (size_t i = 0; i < jb; i += IBs) {
        const size_t cc = j + i, w = std::min(IBs, jb - i);
        double* acc = c.a + cc + cc * c.lda;
        double* inv = c.inv + (cc / IBs) * IBs * IBs;
        NAB_TRY(potf2(sp, acc, c.lda, (int)w, c.use_sub, c.sub, cc + c.col_offset, c.fail, inv));
        const size_t r = n - cc - w;
        if (r == 0) continue;
        double* a21 = acc + w;
        set_gemm_sm_limit(sm_limit);
        // A21 <- A21 * inv(L)^T, in place: the result is one tile column wide (w <= 128 = BN), so the output tile of a
        // CTA is exactly the A operand rows it has finished reading (all its k-blocks have landed in shared memory
        // before the epilogue stores), and no other CTA touches those rows.
        (void)tmp; (void)ldt;
        int st = dgemm_device(sp, false, r, w, w, 1.0, a21, 1, (ptrdiff_t)c.lda, inv, (ptrdiff_t)IBs, 1, 0.0, a21, 1, (ptrdiff_t)c.lda);
        const size_t pw = j + jb - cc - w;               // remaining columns of this panel
        if (st == NA_OK && pw > 0)
            st = dgemm_device(sp, true, r, w, pw, -1.0, a21, 1, (ptrdiff_t)c.lda, a21, (ptrdiff_t)c.lda, 1, 1.0,
                              acc + w + w * c.lda, 1, (ptrdiff_t)c.lda);
        set_gemm_sm_limit(0);
        NAB_TRY(st);
    }
    return NA_OK;
}

static int chol_lookahead(const CholCtx& c, size_t n) {
    StreamGuard su_g;
    EventGuard ev_p, ev_u, ev_d;
    NAB_TRY(su_g.create()); NAB_TRY(ev_p.create()); NAB_TRY(ev_u.create()); NAB_TRY(ev_d.create());
    const cudaStream_t sp = c.s, su = su_g.s;
    Scratch tmp;
    const size_t ldt = round_up(n, 2);
    int st = tmp.alloc(ldt * IBs * sizeof(double), sp);
    const int sms = ctx().sm_count;
    const size_t NB = CHOL_NB;
    bool bulk_pending = false;
    Timeline tr("NAB_CHOL_TRACE", "chol_trace");
    tr.start(sp);
    cudaEvent_t t_first = tr.mark(sp);
    if (st == NA_OK) st = chol_panel(c, sp, n, 0, std::min(NB, n), 0, tmp.as<double>(), ldt);
    tr.add("panel", 0, t_first, tr.mark(sp));
    if (st == NA_OK) st = sink_panel(c, sp, n, 0, std::min(NB, n));
    for (size_t j = 0; st == NA_OK && j + NB < n; j += NB) {
        const size_t jb = NB, jn = j + jb, jbn = std::min(NB, n - jn);
        const double* pj = c.a + j * c.lda;               // panel j: columns [j, j+jb)
        // la(j): next panel's columns, rows jn.., needs bulk(j - nb) finished on those columns
        if (bulk_pending) { cudaStreamWaitEvent(sp, ev_u, 0); }
        cudaEvent_t t_la = tr.mark(sp);
        st = dgemm_device(sp, true, n - jn, jb, jbn, -1.0, pj + jn, 1, (ptrdiff_t)c.lda, pj + jn, (ptrdiff_t)c.lda, 1, 1.0,
                          c.a + jn + jn * c.lda, 1, (ptrdiff_t)c.lda);
        if (st != NA_OK) break;
        tr.add("la", j, t_la, tr.mark(sp));
        cudaEventRecord(ev_p, sp);
        const size_t jr = jn + jbn, rr = n - jr;          // bulk region: rows/cols [jr, n)
        int rp = 0;
        size_t wa = 0;                                    // columns of the bulk region updated while the panel runs
        if (rr > 0) {
            // SMs for the panel chain from a simple time model: panel(r) = 4 leaves of ~100 us + m*w^2 TRSM/SYRK
            // flops at the K = 128 rate on r SMs; bulk at the K = nb rate.  Without the split the step costs
            // max(panel(r), bulk(sms - r)); with it, part A (the leftmost `wa` columns) runs beside the panel on
            // sms - r CTAs and part B takes the whole GPU when the panel is done: panel(r) + rest / sms.
            const bool split = chol_split();
            double t_panel = 0.0;
            {
                const double m_p = (double)(n - jn), w_p = (double)jbn;
                const double bulk_flops = (double)jb * (double)rr * (double)rr;          // 2 * K * rr^2 / 2
                double best = 1e30;
                rp = 16;
                for (int r = chol_rp_min(); r <= sms - 28; r += 4) {
                    const double tp = (w_p / 128.0) * 100e-6 + m_p * w_p * w_p / (r * 0.15e12);
                    const double tb = bulk_flops / ((sms - r) * kSmFlops);
                    const double rest = std::max(0.0, bulk_flops - tp * (sms - r) * kSmFlops);
                    const double t = split ? tp + rest / (sms * kSmFlops) : std::max(tp, tb);
                    if (t < best) { best = t; rp = r; t_panel = tp; }
                }
            }
            if (chol_rp_override() > 0) rp = chol_rp_override();
            wa = rr;
            if (split) {
                const double target = t_panel * (sms - rp) * kSmFlops;
                for (wa = IBs; wa < rr; wa += IBs) {
                    const double area = (double)wa * ((double)rr - 0.5 * (double)wa);
                    if (2.0 * (double)jb * area >= target) break;
                }
                if (wa + 2 * IBs >= rr) wa = rr;
            }
            cudaStreamWaitEvent(su, ev_p, 0);
            cudaEvent_t t_a = tr.mark(su);
            set_gemm_sm_limit(sms - rp);
            st = dgemm_device(su, true, rr, jb, wa, -1.0, pj + jr, 1, (ptrdiff_t)c.lda, pj + jr, (ptrdiff_t)c.lda, 1, 1.0,
                              c.a + jr + jr * c.lda, 1, (ptrdiff_t)c.lda);
            set_gemm_sm_limit(0);
            if (st != NA_OK) break;
            tr.add("bulkA", j, t_a, tr.mark(su));
        }
        cudaEvent_t t_p = tr.mark(sp);
        st = chol_panel(c, sp, n, jn, jbn, rp, tmp.as<double>(), ldt);
        if (st != NA_OK) break;
        tr.add("panel", jn, t_p, tr.mark(sp));
        st = sink_panel(c, sp, n, jn, jbn);
        if (st != NA_OK) break;
        if (tr.on) fprintf(stderr, "chol_trace j=%6zu rp=%d wa=%zu rr=%zu\n", j, rp, wa, rr);
        if (rr > 0) {
            if (wa < rr) {
                cudaEventRecord(ev_d, sp);
                cudaStreamWaitEvent(su, ev_d, 0);
                const size_t jq = jr + wa, rq = n - jq;
                st = dgemm_device(su, true, rq, jb, rq, -1.0, pj + jq, 1, (ptrdiff_t)c.lda, pj + jq, (ptrdiff_t)c.lda, 1, 1.0,
                                  c.a + jq + jq * c.lda, 1, (ptrdiff_t)c.lda);
                if (st != NA_OK) break;
            }
            cudaEventRecord(ev_u, su);
            bulk_pending = true;
        }
    }
    if (bulk_pending) cudaStreamWaitEvent(sp, ev_u, 0);
    cudaStreamSynchronize(su);
    tr.dump();
    return st;
}

// Everything but the status read-back: enqueues the factorization on `s`; the first failing column (offset by
// col_offset) is atomicMin'ed into the device word *fail_dev, which the caller initialised to ~0.
static int cholesky_device_async_sink(cudaStream_t s, size_t n, double* a, size_t lda, int use_sub, double sub,
                                      unsigned long long* fail_dev, size_t col_offset, const HostSink* sink) {
    if (n == 0) return NA_OK;
    if (lda < n) { set_error("cholesky: lda < n"); return NA_EINVAL; }
    Scratch inv;
    NAB_TRY(inv.alloc(ceil_div(n, IBs) * IBs * IBs * sizeof(double), s));
    CholCtx c{s, a, lda, use_sub, sub, inv.as<double>(), fail_dev, col_offset, sink};
    if (n <= 2 * CHOL_NB) NAB_TRY(chol_rec(c, 0, n));
    else NAB_TRY(chol_lookahead(c, n));
    return NA_OK;
}

int cholesky_device_async(cudaStream_t s, size_t n, double* a, size_t lda, int use_sub, double sub,
                          unsigned long long* fail_dev, size_t col_offset) {
    return cholesky_device_async_sink(s, n, a, lda, use_sub, sub, fail_dev, col_offset, nullptr);
}

int cholesky_device(cudaStream_t s, size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col) {
    if (n == 0) return NA_OK;
    if (lda < n) { set_error("cholesky: lda < n"); return NA_EINVAL; }
    Scratch flag;
    NAB_TRY(flag.alloc(sizeof(unsigned long long), s));
    NAB_CUDA(cudaMemsetAsync(flag.p, 0xff, sizeof(unsigned long long), s));
    NAB_TRY(cholesky_device_async(s, n, a, lda, use_sub, sub, flag.as<unsigned long long>(), 0));
    unsigned long long h = 0;
    NAB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    if (h != ~0ull) {
        if (fail_col) *fail_col = (size_t)h;
        return NA_NOT_PD;
    }
    return NA_OK;
}

// Cholesky::solve_mut: L y = b, then L^T x = y (cholesky.rs:127-128).
int cholesky_solve_device(cudaStream_t s, size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs) {
    if (n == 0 || nrhs == 0) return NA_OK;
    Scratch inv;
    NAB_TRY(inv.alloc(ceil_div(n, IBs) * IBs * IBs * sizeof(double), s));
    NAB_TRY(trtri_blocks(s, l, 1, (ptrdiff_t)lda, n, true, false, nullptr, inv.as<double>()));
    NAB_TRY(trsm_left(s, true, false, n, l, 1, (ptrdiff_t)lda, nullptr, inv.as<double>(), b, 1, (ptrdiff_t)ldb, nrhs));
    // L^T is effectively upper; inverse of the diagonal blocks of L^T = transposes of inv(L_bb)
    Scratch invt;
    NAB_TRY(invt.alloc(ceil_div(n, IBs) * IBs * IBs * sizeof(double), s));
    NAB_TRY(trtri_blocks(s, l, (ptrdiff_t)lda, 1, n, false, false, nullptr, invt.as<double>()));
    return trsm_left(s, false, false, n, l, (ptrdiff_t)lda, 1, nullptr, invt.as<double>(), b, 1, (ptrdiff_t)ldb, nrhs);
}

// ---- host <-> device staging of an ld-strided column-major matrix --------------------------------
int upload_matrix(cudaStream_t s, Scratch& buf, size_t& ldd, const double* h, size_t ldh, size_t rows, size_t cols) {
    ldd = round_up(std::max<size_t>(rows, 1), 2);
    NAB_TRY(buf.alloc(ldd * std::max<size_t>(cols, 1) * sizeof(double), s));
    if (rows && cols) NAB_CUDA(cudaMemcpy2DAsync(buf.p, ldd * 8, h, ldh * 8, rows * 8, cols, cudaMemcpyHostToDevice, s));
    return NA_OK;
}
int download_matrix(cudaStream_t s, double* h, size_t ldh, const double* d, size_t ldd, size_t rows, size_t cols) {
    if (rows && cols) NAB_CUDA(cudaMemcpy2DAsync(h, ldh * 8, d, ldd * 8, rows * 8, cols, cudaMemcpyDeviceToHost, s));
    return NA_OK;
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_cholesky_f64_dev(size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col, void* stream) {
    NAB_TRY(ensure_init());
    return cholesky_device(static_cast<cudaStream_t>(stream), n, a, lda, use_sub, sub, fail_col);
}

int na_cholesky_f64_dev_async(size_t n, double* a, size_t lda, int use_sub, double sub, uint64_t* fail_col_dev, size_t col_offset,
                              void* stream) {
    NAB_TRY(ensure_init());
    if (!fail_col_dev) { set_error("cholesky (async): fail_col_dev is null"); return NA_EINVAL; }
    return cholesky_device_async(static_cast<cudaStream_t>(stream), n, a, lda, use_sub, sub,
                                 reinterpret_cast<unsigned long long*>(fail_col_dev), col_offset);
}

int na_cholesky_f64(size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col) {
    NAB_TRY(ensure_init());
    if (n == 0) return NA_OK;
    if (!a || lda < n) { set_error("cholesky: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    if (n <= 2 * CHOL_NB) {           // small: one upload, factor, one download
        Scratch d; size_t ldd;
        NAB_TRY(upload_matrix(s, d, ldd, a, lda, n, n));
        int st = cholesky_device(s, n, d.as<double>(), ldd, use_sub, sub, fail_col);
        if (st < 0) return st;
        NAB_TRY(download_matrix(s, a, lda, d.as<double>(), ldd, n, n));
        NAB_CUDA(cudaStreamSynchronize(s));
        return st;
    }
    // Large: only the lower triangle crosses PCIe (the reference never reads or writes the strict upper one,
    // cholesky.rs:221-272), as column-block trapezoids; finished panels stream back behind the factorization
    // (overlap needs pinned host memory, na_host_alloc_pinned).
    StreamGuard sc;
    EventGuard ev;
    NAB_TRY(sc.create()); NAB_TRY(ev.create());
    Scratch d, flag;
    const size_t ldd = round_up(n, 2), NB = CHOL_NB;
    NAB_TRY(d.alloc(ldd * n * sizeof(double), s));
    NAB_TRY(flag.alloc(sizeof(unsigned long long), s));
    NAB_CUDA(cudaMemsetAsync(flag.p, 0xff, sizeof(unsigned long long), s));
    for (size_t j = 0; j < n; j += NB) {
        const size_t jb = std::min(NB, n - j);
        NAB_CUDA(cudaMemcpy2DAsync(d.as<double>() + j + j * ldd, ldd * 8, a + j + j * lda, lda * 8, (n - j) * 8, jb, cudaMemcpyHostToDevice, s));
    }
    HostSink sink{a, lda, sc.s, ev.e};
    {
        const int st = cholesky_device_async_sink(s, n, d.as<double>(), ldd, use_sub, sub, flag.as<unsigned long long>(), 0, &sink);
        if (st < 0) { cudaStreamSynchronize(s); cudaStreamSynchronize(sc.s); return st; }    // copies still read `d`
    }
    unsigned long long h = 0;
    NAB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    NAB_CUDA(cudaStreamSynchronize(sc.s));
    if (h != ~0ull) {
        if (fail_col) *fail_col = (size_t)h;
        return NA_NOT_PD;
    }
    return NA_OK;
}

int na_cholesky_solve_f64_dev(size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs, void* stream) {
    NAB_TRY(ensure_init());
    return cholesky_solve_device(static_cast<cudaStream_t>(stream), n, l, lda, b, ldb, nrhs);
}

int na_cholesky_solve_f64(size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs) {
    NAB_TRY(ensure_init());
    if (n == 0 || nrhs == 0) return NA_OK;
    if (!l || !b || lda < n || ldb < n) { set_error("cholesky_solve: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch dl, db; size_t ldl, lddb;
    NAB_TRY(upload_matrix(s, dl, ldl, l, lda, n, n));
    NAB_TRY(upload_matrix(s, db, lddb, b, ldb, n, nrhs));
    NAB_TRY(cholesky_solve_device(s, n, dl.as<double>(), ldl, db.as<double>(), lddb, nrhs));
    NAB_TRY(download_matrix(s, b, ldb, db.as<double>(), lddb, n, nrhs));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

// op(T) x = b.  NA_SINGULAR on an exactly-zero diagonal (checked variants of solve.rs:55-182).
int na_tri_solve_f64_dev(int lower, int trans, int unit_diag, size_t n, const double* t, size_t ldt,
                         double* b, size_t ldb, size_t nrhs, void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n == 0 || nrhs == 0) return NA_OK;
    if (!t || !b || ldt < n || ldb < n) { set_error("tri_solve: bad arguments"); return NA_EINVAL; }
    if (!unit_diag) {
        Scratch flag;
        NAB_TRY(flag.alloc(sizeof(int), s));
        NAB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
        NAB_TRY(zero_diag_check(s, t, ldt, nullptr, n, flag.as<int>()));
        int h = 0;
        NAB_CUDA(cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
        if (h) return NA_SINGULAR;
    }
    const bool eff_lower = (lower != 0) != (trans != 0);
    const ptrdiff_t rsm = trans ? (ptrdiff_t)ldt : 1, csm = trans ? 1 : (ptrdiff_t)ldt;
    return trsm_left(s, eff_lower, unit_diag != 0, n, t, rsm, csm, nullptr, nullptr, b, 1, (ptrdiff_t)ldb, nrhs);
}

int na_trsm_f64_dev(int side_right, int lower, int trans, int unit_diag, size_t m, size_t n,
                    const double* t, size_t ldt, double* b, size_t ldb, void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (m == 0 || n == 0) return NA_OK;
    const size_t nt = side_right ? n : m;
    if (!t || !b || ldt < nt || ldb < m) { set_error("trsm: bad arguments"); return NA_EINVAL; }
    if (!side_right) {
        // unit-lower, untransposed, at most 128 rows: the direct substitution kernel of the LU panels (one launch,
        // no inverse blocks)
        if (lower && !trans && unit_diag && m <= 128) return trsm_unit_lower_small(s, m, t, ldt, b, ldb, n);
        const bool eff_lower = (lower != 0) != (trans != 0);
        const ptrdiff_t rsm = trans ? (ptrdiff_t)ldt : 1, csm = trans ? 1 : (ptrdiff_t)ldt;
        return trsm_left(s, eff_lower, unit_diag != 0, m, t, rsm, csm, nullptr, nullptr, b, 1, (ptrdiff_t)ldb, n);
    }
    // X op(T) = B  <=>  op(T)^T X^T = B^T: a left solve on the transposed view of B
    const bool mt = trans == 0;                         // M = op(T)^T is T^T when op = T, and T when op = T^T
    const bool eff_lower = (lower != 0) != mt;
    const ptrdiff_t rsm = mt ? (ptrdiff_t)ldt : 1, csm = mt ? 1 : (ptrdiff_t)ldt;
    return trsm_left(s, eff_lower, unit_diag != 0, n, t, rsm, csm, nullptr, nullptr, b, (ptrdiff_t)ldb, 1, m);
}

int na_dgemm_lower_dev(size_t m, size_t k, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                       const double* b, ptrdiff_t rsb, ptrdiff_t csb, double beta, double* c, size_t ldc, void* stream) {
    NAB_TRY(ensure_init());
    if (m < n) { set_error("gemm_lower: needs m >= n"); return NA_EINVAL; }
    if (k == 0) { set_error("gemm_lower: k == 0 is not supported"); return NA_EINVAL; }
    return dgemm_device(static_cast<cudaStream_t>(stream), true, m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, 1, (ptrdiff_t)ldc);
}

int na_tri_solve_f64(int lower, int trans, int unit_diag, size_t n, const double* t, size_t ldt,
                     double* b, size_t ldb, size_t nrhs) {
    NAB_TRY(ensure_init());
    if (n == 0 || nrhs == 0) return NA_OK;
    if (!t || !b || ldt < n || ldb < n) { set_error("tri_solve: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch dt, db; size_t ldd, lddb;
    NAB_TRY(upload_matrix(s, dt, ldd, t, ldt, n, n));
    NAB_TRY(upload_matrix(s, db, lddb, b, ldb, n, nrhs));
    int st = na_tri_solve_f64_dev(lower, trans, unit_diag, n, dt.as<double>(), ldd, db.as<double>(), lddb, nrhs, s);
    if (st != NA_OK) return st;
    NAB_TRY(download_matrix(s, b, ldb, db.as<double>(), lddb, n, nrhs));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

}  // extern "C"
