// capi_gemm.cu -- extern "C" GEMM entry points (seam 1: matrixmultiply::dgemm / sgemm).
#include "common.cuh"
#include "kernels.cuh"

#include <utility>
#include <vector>

using namespace nab;

namespace nab {

// A host matrix staged on the device: column-major `rows x cols` (ld even, 256B-aligned base),
// possibly holding the TRANSPOSE of the host view when the host view is row-major (then the
// device strides handed to the GEMM are swapped instead of moving data twice).
struct Staged {
    Scratch buf;
    size_t ld = 0;
    ptrdiff_t rs = 1, cs = 0;           // device element strides of the logical (rows x cols) view
    std::vector<double> host_tmp;       // gather buffer for general strides
};

static ptrdiff_t pabs(ptrdiff_t x) { return x < 0 ? -x : x; }

// Uploads (when `upload`) the logical rows x cols host view (rs, cs) and fills in device strides.
static int stage_in(cudaStream_t s, Staged& st, const double* h, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols, bool upload) {
    if (rs == 1 && (cs >= (ptrdiff_t)rows || cols == 1)) {                 // column-major view
        st.ld = round_up(rows, 2); st.rs = 1; st.cs = (ptrdiff_t)st.ld;
        NAB_TRY(st.buf.alloc(st.ld * cols * sizeof(double), s));
        if (upload) NAB_CUDA(cudaMemcpy2DAsync(st.buf.p, st.ld * 8, h, (cols == 1 ? rows : (size_t)cs) * 8, rows * 8, cols, cudaMemcpyHostToDevice, s));
        return NA_OK;
    }
    if (cs == 1 && (rs >= (ptrdiff_t)cols || rows == 1)) {                 // row-major view: stage its transpose
        st.ld = round_up(cols, 2); st.rs = (ptrdiff_t)st.ld; st.cs = 1;
        NAB_TRY(st.buf.alloc(st.ld * rows * sizeof(double), s));
        if (upload) NAB_CUDA(cudaMemcpy2DAsync(st.buf.p, st.ld * 8, h, (rows == 1 ? cols : (size_t)rs) * 8, cols * 8, rows, cudaMemcpyHostToDevice, s));
        return NA_OK;
    }
    // general strides (both != 1, or negative): gather on the host
    st.ld = round_up(rows, 2); st.rs = 1; st.cs = (ptrdiff_t)st.ld;
    NAB_TRY(st.buf.alloc(st.ld * cols * sizeof(double), s));
    if (upload) {
        st.host_tmp.resize(rows * cols);
        for (size_t j = 0; j < cols; ++j)
            for (size_t i = 0; i < rows; ++i) st.host_tmp[i + j * rows] = h[(ptrdiff_t)i * rs + (ptrdiff_t)j * cs];
        NAB_CUDA(cudaMemcpy2DAsync(st.buf.p, st.ld * 8, st.host_tmp.data(), rows * 8, rows * 8, cols, cudaMemcpyHostToDevice, s));
    }
    return NA_OK;
}

// Downloads the staged rows x cols result back into the host view. Synchronises the stream.
static int stage_out(cudaStream_t s, Staged& st, double* h, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols) {
    if (st.rs == 1 && rs == 1 && (cs >= (ptrdiff_t)rows || cols == 1)) {
        NAB_CUDA(cudaMemcpy2DAsync(h, (cols == 1 ? rows : (size_t)cs) * 8, st.buf.p, st.ld * 8, rows * 8, cols, cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
        return NA_OK;
    }
    if (st.cs == 1 && cs == 1) {
        NAB_CUDA(cudaMemcpy2DAsync(h, (rows == 1 ? cols : (size_t)rs) * 8, st.buf.p, st.ld * 8, cols * 8, rows, cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
        return NA_OK;
    }
    st.host_tmp.resize(rows * cols);
    NAB_CUDA(cudaMemcpy2DAsync(st.host_tmp.data(), rows * 8, st.buf.p, st.ld * 8, rows * 8, cols, cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    for (size_t j = 0; j < cols; ++j)
        for (size_t i = 0; i < rows; ++i) h[(ptrdiff_t)i * rs + (ptrdiff_t)j * cs] = st.host_tmp[i + j * rows];
    return NA_OK;
}

static int check_view(const char* name, const void* p, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return NA_OK;
    if (!p) { set_error("gemm: %s is null", name); return NA_EINVAL; }
    if ((rows > 1 && rs == 0) || (cols > 1 && cs == 0)) {
        // zero strides are legal for *inputs* in matrixmultiply (broadcast); we keep them legal too
        return NA_OK;
    }
    (void)pabs;
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Pipelined host-pointer GEMM for large column-major operands: C is cut into a gr x gc grid of
// chunks visited shell by shell, so that each chunk needs at most one new row block of A or column
// block of B and the demand for new panels is spread over the whole run.  H2D copies (stream h2d), chunk GEMMs (stream cmp) and D2H copies of finished chunks
// (stream d2h) overlap; only the first A/B pieces and the last C chunk are exposed.
// ------------------------------------------------------------------------------------------------
static int dgemm_host_pipelined(size_t m, size_t k, size_t n, double alpha, const double* a, size_t lda,
                                const double* b, size_t ldb, double beta, double* c, size_t ldc) {
    Context& cx = ctx();
    cudaStream_t cmp = cx.stream, h2d = cx.stream2, d2h = nullptr;
    NAB_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    // Non-uniform chunk grid: a large first block (half of the rows / columns), then quarters.  The first chunk is
    // the one that has to hide the upload of two panels behind its own compute -- (r + c) K bytes against 2 r c K
    // flops, so it must be big (measured with a uniform 4 x 4 grid: the first two chunks waited 14 ms for PCIe) --
    // and the last chunk, whose D2H copy is exposed, must be small.
    auto cuts = [](size_t len) {
        std::vector<size_t> v{0};
        if (len >= 8192) { v.push_back(round_up(len / 2, 128)); v.push_back(round_up(3 * len / 4, 128)); }
        else v.push_back(round_up(len / 2, 128));
        v.push_back(len);
        return v;
    };
    const std::vector<size_t> rb = cuts(m), cb = cuts(n);
    const size_t nbr = rb.size() - 1, nbc = cb.size() - 1;
    Scratch da, db, dc;
    const size_t ldda = round_up(m, 2), lddb = round_up(k, 2), lddc = round_up(m, 2);
    int st = da.alloc(ldda * k * 8, cmp);
    if (st == NA_OK) st = db.alloc(lddb * n * 8, cmp);
    if (st == NA_OK) st = dc.alloc(lddc * n * 8, cmp);
    std::vector<cudaEvent_t> ev_a(nbr, nullptr), ev_b(nbc, nullptr), ev_c(nbr * nbc, nullptr), ev_cin(nbr * nbc, nullptr);
    cudaEvent_t ev_alloc = nullptr;
    auto mk = [](cudaEvent_t& e) { return cudaEventCreateWithFlags(&e, cudaEventDisableTiming); };
    if (st == NA_OK) {
        mk(ev_alloc); cudaEventRecord(ev_alloc, cmp);
        cudaStreamWaitEvent(h2d, ev_alloc, 0); cudaStreamWaitEvent(d2h, ev_alloc, 0);
    }
    Timeline tr("NAB_GEMM_TRACE", "gemm_trace");
    tr.start(cmp);
    std::vector<bool> a_up(nbr, false), b_up(nbc, false);
    // The first chunk is additionally cut along K: its A row block and B column block arrive as `ks` K-slabs (>= 1024
    // deep), and the chunk is computed as ks accumulating GEMMs, so the compute stream starts after 1/ks of the first
    // two panels.
    const size_t ks = k >= 2048 ? std::min<size_t>(16, k / 1024) : 1;
    if (st == NA_OK && ks > 1) {
        const size_t rows = rb[1], cols = cb[1];
        const size_t kw = round_up(ceil_div(k, ks), 16);
        double* dcc = dc.as<double>();
        cudaEvent_t ev_c0 = nullptr;
        if (beta != 0.0) {
            if (cudaMemcpy2DAsync(dcc, lddc * 8, c, ldc * 8, rows * 8, cols, cudaMemcpyHostToDevice, h2d) != cudaSuccess) st = NA_ECUDA;
            mk(ev_c0); cudaEventRecord(ev_c0, h2d); cudaStreamWaitEvent(cmp, ev_c0, 0);
        }
        std::vector<cudaEvent_t> ev_s;
        for (size_t k0 = 0; st == NA_OK && k0 < k; k0 += kw) {
            const size_t kk = std::min(kw, k - k0);
            if (cudaMemcpy2DAsync(da.as<double>() + k0 * ldda, ldda * 8, a + k0 * lda, lda * 8, rows * 8, kk, cudaMemcpyHostToDevice, h2d) != cudaSuccess ||
                cudaMemcpy2DAsync(db.as<double>() + k0, lddb * 8, b + k0, ldb * 8, kk * 8, cols, cudaMemcpyHostToDevice, h2d) != cudaSuccess) { st = NA_ECUDA; break; }
            cudaEvent_t e = nullptr; mk(e); cudaEventRecord(e, h2d); ev_s.push_back(e);
            cudaStreamWaitEvent(cmp, e, 0);
            cudaEvent_t tg = tr.mark(cmp);
            st = dgemm_device(cmp, false, rows, kk, cols, alpha, da.as<double>() + k0 * ldda, 1, (ptrdiff_t)ldda,
                              db.as<double>() + k0, 1, (ptrdiff_t)lddb, k0 == 0 ? beta : 1.0, dcc, 1, (ptrdiff_t)lddc);
            tr.add("slab", k0, tg, tr.mark(cmp));
        }
        if (st == NA_OK) {
            a_up[0] = true; b_up[0] = true;
            mk(ev_a[0]); cudaEventRecord(ev_a[0], h2d); mk(ev_b[0]); cudaEventRecord(ev_b[0], h2d);
            mk(ev_c[0]); cudaEventRecord(ev_c[0], cmp);
            cudaStreamWaitEvent(d2h, ev_c[0], 0);
            if (cudaMemcpy2DAsync(c, ldc * 8, dcc, lddc * 8, rows * 8, cols, cudaMemcpyDeviceToHost, d2h) != cudaSuccess) st = NA_ECUDA;
        }
        for (cudaEvent_t e : ev_s) ev_cin.push_back(e);      // destroyed with the others once the streams are idle
        if (ev_c0) ev_cin.push_back(ev_c0);
    }
    // Chunk order: shells.  Shell L adds column block L (chunks (0..L-1, L): only B_L is new), then the corner (L, L)
    // (A_L is new), then row block L (chunks (L, L-1..0): nothing new).  New panels are needed at steps 0, 1, 2, 4, 6,
    // 9, 12 of 16 instead of at every step of the first row, so the H2D stream stays ahead of the compute stream.
    std::vector<std::pair<size_t, size_t>> order;
    for (size_t L = 0; L < std::max(nbr, nbc); ++L) {
        const bool last_shell = L + 1 == std::max(nbr, nbc);      // its corner (the smallest chunk) goes last: its D2H is exposed
        if (L < nbc) for (size_t i = 0; i < std::min(L, nbr); ++i) order.push_back({i, L});
        if (L < nbr && L < nbc && !last_shell) order.push_back({L, L});
        if (L < nbr) for (size_t j = std::min(L, nbc); j-- > 0;) order.push_back({L, j});
        if (L < nbr && L < nbc && last_shell) order.push_back({L, L});
    }
    for (size_t step = (ks > 1 ? 1 : 0); st == NA_OK && step < order.size(); ++step) {
        const size_t bi = order[step].first, bj = order[step].second;
        const size_t r0 = rb[bi], rows = rb[bi + 1] - r0, c0 = cb[bj], cols = cb[bj + 1] - c0;
        if (!a_up[bi]) {
            if (cudaMemcpy2DAsync(da.as<double>() + r0, ldda * 8, a + r0, lda * 8, rows * 8, k, cudaMemcpyHostToDevice, h2d) != cudaSuccess) { st = NA_ECUDA; break; }
            mk(ev_a[bi]); cudaEventRecord(ev_a[bi], h2d); a_up[bi] = true;
        }
        if (!b_up[bj]) {
            if (cudaMemcpy2DAsync(db.as<double>() + c0 * lddb, lddb * 8, b + c0 * ldb, ldb * 8, k * 8, cols, cudaMemcpyHostToDevice, h2d) != cudaSuccess) { st = NA_ECUDA; break; }
            mk(ev_b[bj]); cudaEventRecord(ev_b[bj], h2d); b_up[bj] = true;
        }
        double* dcc = dc.as<double>() + r0 + c0 * lddc;
        if (beta != 0.0) {      // C is only read (and therefore only uploaded) when beta != 0
            if (cudaMemcpy2DAsync(dcc, lddc * 8, c + r0 + c0 * ldc, ldc * 8, rows * 8, cols, cudaMemcpyHostToDevice, h2d) != cudaSuccess) { st = NA_ECUDA; break; }
            mk(ev_cin[step]); cudaEventRecord(ev_cin[step], h2d);
            cudaStreamWaitEvent(cmp, ev_cin[step], 0);
        }
        cudaStreamWaitEvent(cmp, ev_a[bi], 0);
        cudaStreamWaitEvent(cmp, ev_b[bj], 0);
        cudaEvent_t tg = tr.mark(cmp);
        st = dgemm_device(cmp, false, rows, k, cols, alpha, da.as<double>() + r0, 1, (ptrdiff_t)ldda,
                          db.as<double>() + c0 * lddb, 1, (ptrdiff_t)lddb, beta, dcc, 1, (ptrdiff_t)lddc);
        if (st != NA_OK) break;
        tr.add("chunk", step, tg, tr.mark(cmp));
        mk(ev_c[step]); cudaEventRecord(ev_c[step], cmp);
        cudaStreamWaitEvent(d2h, ev_c[step], 0);
        cudaEvent_t td = tr.mark(d2h);
        if (cudaMemcpy2DAsync(c + r0 + c0 * ldc, ldc * 8, dcc, lddc * 8, rows * 8, cols, cudaMemcpyDeviceToHost, d2h) != cudaSuccess) { st = NA_ECUDA; break; }
        tr.add("d2h", step, td, tr.mark(d2h));
    }
    cudaError_t e1 = cudaStreamSynchronize(d2h), e2 = cudaStreamSynchronize(h2d), e3 = cudaStreamSynchronize(cmp);
    tr.dump();
    for (auto* v : {&ev_a, &ev_b, &ev_c, &ev_cin}) for (cudaEvent_t e : *v) if (e) cudaEventDestroy(e);
    if (ev_alloc) cudaEventDestroy(ev_alloc);
    cudaStreamDestroy(d2h);
    if (st != NA_OK) { if (st == NA_ECUDA) set_error("pipelined gemm: CUDA copy failed (%s)", cudaGetErrorString(cudaGetLastError())); return st; }
    if (e1 != cudaSuccess) return cuda_fail(e1, "pipelined gemm d2h", __FILE__, __LINE__);
    if (e2 != cudaSuccess) return cuda_fail(e2, "pipelined gemm h2d", __FILE__, __LINE__);
    if (e3 != cudaSuccess) return cuda_fail(e3, "pipelined gemm compute", __FILE__, __LINE__);
    return NA_OK;
}

}  // namespace nab

extern "C" {

int na_dgemm_dev(size_t m, size_t k, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const double* b, ptrdiff_t rsb, ptrdiff_t csb, double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc,
                 void* stream) {
    NAB_TRY(ensure_init());
    return dgemm_device(static_cast<cudaStream_t>(stream), false, m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc);
}

int na_dgemm(size_t m, size_t k, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
             const double* b, ptrdiff_t rsb, ptrdiff_t csb, double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc) {
    NAB_TRY(ensure_init());
    if (m == 0 || n == 0) return NA_OK;
    NAB_TRY(check_view("a", a, rsa, csa, m, k));
    NAB_TRY(check_view("b", b, rsb, csb, k, n));
    NAB_TRY(check_view("c", c, rsc, csc, m, n));
    if ((m > 1 && rsc == 0) || (n > 1 && csc == 0)) { set_error("gemm: c has a zero stride"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    if (k == 0) {   // blas_uninit.rs:258-269: C <- beta*C or zeros, no device round trip needed for the semantics,
                    // but the product has no CPU compute path: do it on the device like everything else.
        Staged sc;
        NAB_TRY(stage_in(s, sc, c, rsc, csc, m, n, beta != 0.0));
        NAB_TRY(scale_strided(s, sc.buf.as<double>(), sc.rs, sc.cs, m, n, beta));
        return stage_out(s, sc, c, rsc, csc, m, n);
    }
    // large plain column-major operands (VecStorage): overlap the PCIe copies with the compute
    if (rsa == 1 && rsb == 1 && rsc == 1 && m >= 2048 && n >= 2048 && k >= 512 &&
        csa >= (ptrdiff_t)m && csb >= (ptrdiff_t)k && csc >= (ptrdiff_t)m)
        return dgemm_host_pipelined(m, k, n, alpha, a, (size_t)csa, b, (size_t)csb, beta, c, (size_t)csc);
    Staged sa, sb, sc;
    NAB_TRY(stage_in(s, sa, a, rsa, csa, m, k, true));
    NAB_TRY(stage_in(s, sb, b, rsb, csb, k, n, true));
    NAB_TRY(stage_in(s, sc, c, rsc, csc, m, n, beta != 0.0));   // C is not read when beta == 0
    NAB_TRY(dgemm_device(s, false, m, k, n, alpha, sa.buf.as<double>(), sa.rs, sa.cs, sb.buf.as<double>(), sb.rs, sb.cs,
                         beta, sc.buf.as<double>(), sc.rs, sc.cs));
    return stage_out(s, sc, c, rsc, csc, m, n);
}

int na_fill_uniform_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed, void* stream) {
    NAB_TRY(ensure_init());
    return fill_uniform(static_cast<cudaStream_t>(stream), a, nrows, ncols, lda, seed, 0, 0, nrows);
}

int na_fill_uniform_block_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                              size_t row0, size_t col0, size_t global_rows, void* stream) {
    NAB_TRY(ensure_init());
    return fill_uniform(static_cast<cudaStream_t>(stream), a, nrows, ncols, lda, seed, row0, col0, global_rows);
}

int na_set_gemm_sm_limit(int max_ctas) {
    if (max_ctas < 0) { set_error("na_set_gemm_sm_limit: negative limit"); return NA_EINVAL; }
    set_user_gemm_sm_limit(max_ctas);
    return NA_OK;
}

// C (n x n, column-major, ldc) <- alpha * A * A^T + beta * C on the LOWER triangle only; the strict upper triangle of
// the host C is neither read nor written.  Staging: the 128-column block trapezoids [j, n) x [j, j+128) travel each
// way, so the strict upper part only round-trips inside the 128 x 128 diagonal blocks (uploaded first, so what comes
// back there is what was sent).
int na_dsyrk_lower(size_t n, size_t k, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa, double beta, double* c, size_t ldc) {
    NAB_TRY(ensure_init());
    if (n == 0) return NA_OK;
    if (!c || ldc < n || (k && !a)) { set_error("syrk: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Staged sa;
    Scratch dc;
    const size_t ldd = round_up(n, 2), B = 128;
    NAB_TRY(dc.alloc(ldd * n * sizeof(double), s));
    for (size_t j = 0; j < n; j += B) {           // diagonal blocks always (their strict upper part must come back unchanged)
        const size_t w = std::min(B, n - j), rows = beta != 0.0 ? n - j : w;
        NAB_CUDA(cudaMemcpy2DAsync(dc.as<double>() + j + j * ldd, ldd * 8, c + j + j * ldc, ldc * 8, rows * 8, w, cudaMemcpyHostToDevice, s));
    }
    if (k == 0) {
        for (size_t j = 0; j < n; j += B) {       // C <- beta * C (or zeros) on the lower trapezoids, blas_uninit.rs:258-269
            const size_t w = std::min(B, n - j);
            if (n - j > w) NAB_TRY(scale_strided(s, dc.as<double>() + (j + w) + j * ldd, 1, (ptrdiff_t)ldd, n - j - w, w, beta));
            NAB_TRY(scale_lower_block(s, dc.as<double>() + j + j * ldd, ldd, w, beta));
        }
    } else {
        NAB_TRY(stage_in(s, sa, a, rsa, csa, n, k, true));
        NAB_TRY(dgemm_device(s, true, n, k, n, alpha, sa.buf.as<double>(), sa.rs, sa.cs, sa.buf.as<double>(), sa.cs, sa.rs, beta,
                             dc.as<double>(), 1, (ptrdiff_t)ldd));
    }
    for (size_t j = 0; j < n; j += B) {
        const size_t w = std::min(B, n - j);
        NAB_CUDA(cudaMemcpy2DAsync(c + j + j * ldc, ldc * 8, dc.as<double>() + j + j * ldd, ldd * 8, (n - j) * 8, w, cudaMemcpyDeviceToHost, s));
    }
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

// Tuning / diagnostic switches (not needed for correctness).  Keys: "lu_lookahead" (0 = plain recursive path).
int na_set_tuning(const char* key, long value) {
    if (!key) { set_error("na_set_tuning: null key"); return NA_EINVAL; }
    if (strcmp(key, "lu_lookahead") == 0) { lu_set_lookahead(value); return NA_OK; }
    if (strcmp(key, "qr_reg_leaf") == 0) { qr_set_tuning(0, value); return NA_OK; }
    if (strcmp(key, "qr_fused") == 0) { qr_set_tuning(1, value); return NA_OK; }
    if (strcmp(key, "ts_fused") == 0) { ts_set_fused(value); return NA_OK; }
    set_error("na_set_tuning: unknown key '%s'", key);
    return NA_EINVAL;
}

int na_fill_spd_block_dev(double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                          size_t row0, size_t col0, size_t n, void* stream) {
    NAB_TRY(ensure_init());
    return fill_spd(static_cast<cudaStream_t>(stream), a, nrows, ncols, lda, seed, row0, col0, n);
}

}  // extern "C"
