// capi_gemm.cu -- extern "C" GEMM entry points (seam 1: matrixmultiply::dgemm / sgemm).
#include "common.cuh"
#include "kernels.cuh"

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

}  // extern "C"
