// level2.cu -- the Level-1/2 fallbacks of the GEMM path (G8 of SURVEY.md §8a): gemv and axcpy.
//
// Reference: gemv_uninit (/root/reference/src/base/blas_uninit.rs:127-177: y = alpha*A*x + beta*y as a sequence of
// axcpy column updates; y is not read when beta == 0; an empty A scales or zeroes y, :152-160), gemv_tr (blas.rs:
// 503-540: dot products per column) and axcpy (blas_uninit.rs:86-117: y = a*x*c + b*y).  These are what gemm_uninit
// falls back to for small or non-Dyn shapes and what every Level-1 call inside the reference factorizations uses.
// HBM-bound kernels: A is read exactly once, coalesced along its unit stride.
#include "common.cuh"
#include "kernels.cuh"

namespace nab {

// y_part[chunk][i] = sum_{j in chunk} A[i, j] * x[j]   (A column-major-like: unit stride along i)
__global__ void __launch_bounds__(256) gemv_n_partial_kernel(const double* __restrict__ a, long long cs, long long m, long long n,
                                                             const double* __restrict__ x, long long incx, double* __restrict__ part,
                                                             long long cols_per_chunk) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long j0 = (long long)blockIdx.y * cols_per_chunk, j1 = min(n, j0 + cols_per_chunk);
    __shared__ double xs[256];
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    for (long long jb = j0; jb < j1; jb += 256) {
        const long long nj = min((long long)256, j1 - jb);
        __syncthreads();
        if (threadIdx.x < nj) xs[threadIdx.x] = x[(jb + threadIdx.x) * incx];
        __syncthreads();
        if (i < m) {
            const double* ap = a + i + jb * cs;
            long long j = 0;
            for (; j + 4 <= nj; j += 4) {
                acc0 = fma(ap[(j + 0) * cs], xs[j + 0], acc0);
                acc1 = fma(ap[(j + 1) * cs], xs[j + 1], acc1);
                acc2 = fma(ap[(j + 2) * cs], xs[j + 2], acc2);
                acc3 = fma(ap[(j + 3) * cs], xs[j + 3], acc3);
            }
            for (; j < nj; ++j) acc0 = fma(ap[j * cs], xs[j], acc0);
        }
    }
    if (i < m) part[(long long)blockIdx.y * m + i] = (acc0 + acc1) + (acc2 + acc3);
}

// y[i] = alpha * sum_chunks part[chunk][i] + beta * y[i]   (y not read when beta == 0); chunks == 0: y = beta*y / 0
__global__ void gemv_finish_kernel(double* __restrict__ y, long long incy, long long m, const double* __restrict__ part, int chunks,
                                   double alpha, double beta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    for (int c = 0; c < chunks; ++c) s += part[(long long)c * m + i];      // fixed order: deterministic
    double v = alpha * s;
    if (beta != 0.0) v += beta * y[i * incy];
    y[i * incy] = v;
}

// y[j] = alpha * dot(A[:, j], x) + beta * y[j]: one warp per column, lanes along the unit stride
__global__ void __launch_bounds__(256) gemv_t_kernel(const double* __restrict__ a, long long cs, long long m, long long n,
                                                     const double* __restrict__ x, long long incx, double* __restrict__ y, long long incy,
                                                     double alpha, double beta) {
    const long long j = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* ap = a + j * cs;
    double acc0 = 0.0, acc1 = 0.0;
    long long i = lane;
    for (; i + 32 < m; i += 64) { acc0 = fma(ap[i], x[i * incx], acc0); acc1 = fma(ap[i + 32], x[(i + 32) * incx], acc1); }
    if (i < m) acc0 = fma(ap[i], x[i * incx], acc0);
    double s = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        double v = alpha * s;
        if (beta != 0.0) v += beta * y[j * incy];
        y[j * incy] = v;
    }
}

// fully general strides (both != 1): one thread per output, strided reads (rare: views of views)
__global__ void gemv_generic_kernel(const double* __restrict__ a, long long rs, long long cs, long long m, long long n,
                                    const double* __restrict__ x, long long incx, double* __restrict__ y, long long incy, double alpha, double beta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    for (long long j = 0; j < n; ++j) s = fma(a[i * rs + j * cs], x[j * incx], s);
    double v = alpha * s;
    if (beta != 0.0) v += beta * y[i * incy];
    y[i * incy] = v;
}

// y = a * x * c + b * y  (axcpy, blas_uninit.rs:86-117; y not read when b == 0)
__global__ void axcpy_kernel(double* __restrict__ y, long long incy, long long n, double a, const double* __restrict__ x, long long incx,
                             double c, double b) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double t = __dmul_rn(__dmul_rn(a, x[i * incx]), c);          // (a * x) * c, unfused like the reference
        y[i * incy] = b != 0.0 ? __dadd_rn(t, __dmul_rn(b, y[i * incy])) : t;
    }
}

// y (len m) <- alpha * A (m x n, strides rs/cs) * x (len n) + beta * y on device pointers.
int gemv_device(cudaStream_t s, size_t m, size_t n, double alpha, const double* a, ptrdiff_t rs, ptrdiff_t cs,
                const double* x, ptrdiff_t incx, double beta, double* y, ptrdiff_t incy) {
    if (m == 0) return NA_OK;
    if (n == 0) {                                        // blas_uninit.rs:152-160
        gemv_finish_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(y, incy, (long long)m, nullptr, 0, 0.0, beta);
        NAB_LAUNCH_CHECK();
        return NA_OK;
    }
    if (rs == 1 || m == 1) {
        const int sms = ctx().sm_count;
        const size_t row_blocks = ceil_div(m, 256);
        size_t chunks = std::max<size_t>(1, std::min<size_t>(ceil_div((size_t)4 * sms, row_blocks), ceil_div(n, 256)));
        const size_t cpc = round_up(ceil_div(n, chunks), 256);
        chunks = ceil_div(n, cpc);
        Scratch part;
        NAB_TRY(part.alloc(chunks * m * sizeof(double), s));
        gemv_n_partial_kernel<<<dim3((unsigned)row_blocks, (unsigned)chunks), 256, 0, s>>>(a, m == 1 ? (long long)cs : (long long)cs, (long long)m, (long long)n, x,
                                                                                         incx, part.as<double>(), (long long)cpc);
        NAB_LAUNCH_CHECK();
        gemv_finish_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(y, incy, (long long)m, part.as<double>(), (int)chunks, alpha, beta);
        NAB_LAUNCH_CHECK();
        return NA_OK;
    }
    if (cs == 1 || n == 1) {                             // unit stride along the summed index: A x = (A^T)^T x, dot per row
        gemv_t_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, s>>>(a, (long long)rs, (long long)n, (long long)m, x, incx, y, incy, alpha, beta);
        NAB_LAUNCH_CHECK();
        return NA_OK;
    }
    gemv_generic_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, s>>>(a, rs, cs, (long long)m, (long long)n, x, incx, y, incy, alpha, beta);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_dgemv_dev(size_t m, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const double* x, ptrdiff_t incx, double beta, double* y, ptrdiff_t incy, void* stream) {
    NAB_TRY(ensure_init());
    if (m && (!y || (n && (!a || !x)))) { set_error("gemv: null pointer"); return NA_EINVAL; }
    return gemv_device(static_cast<cudaStream_t>(stream), m, n, alpha, a, rsa, csa, x, incx, beta, y, incy);
}

int na_daxcpy_dev(size_t n, double a, const double* x, ptrdiff_t incx, double c, double b, double* y, ptrdiff_t incy, void* stream) {
    NAB_TRY(ensure_init());
    if (n == 0) return NA_OK;
    if (!x || !y) { set_error("axcpy: null pointer"); return NA_EINVAL; }
    const int blocks = (int)std::min<size_t>(ceil_div(n, 256), (size_t)ctx().sm_count * 8);
    axcpy_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, incy, (long long)n, a, x, incx, c, b);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// Host-pointer gemv: stages A (column-major or row-major view; other strides are gathered on the host), x and y.
int na_dgemv(size_t m, size_t n, double alpha, const double* a, ptrdiff_t rsa, ptrdiff_t csa,
             const double* x, ptrdiff_t incx, double beta, double* y, ptrdiff_t incy) {
    NAB_TRY(ensure_init());
    if (m == 0) return NA_OK;
    if (!y || (n && (!a || !x)) || incy == 0) { set_error("gemv: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    std::vector<double> hx(n), hy(m), ha;
    for (size_t j = 0; j < n; ++j) hx[j] = x[(ptrdiff_t)j * incx];
    if (beta != 0.0) for (size_t i = 0; i < m; ++i) hy[i] = y[(ptrdiff_t)i * incy];
    Scratch da, dx, dy;
    size_t ld = round_up(std::max<size_t>(m, 1), 2);
    ptrdiff_t drs = 1, dcs = (ptrdiff_t)ld;
    NAB_TRY(dx.alloc(std::max<size_t>(n, 1) * 8, s));
    NAB_TRY(dy.alloc(m * 8, s));
    if (n) {
        if (rsa == 1 && (csa >= (ptrdiff_t)m || n == 1)) {
            NAB_TRY(da.alloc(ld * n * 8, s));
            NAB_CUDA(cudaMemcpy2DAsync(da.p, ld * 8, a, (n == 1 ? m : (size_t)csa) * 8, m * 8, n, cudaMemcpyHostToDevice, s));
        } else if (csa == 1 && (rsa >= (ptrdiff_t)n || m == 1)) {        // row-major view: stage as is, swap the strides
            ld = round_up(n, 2); drs = (ptrdiff_t)ld; dcs = 1;
            NAB_TRY(da.alloc(ld * m * 8, s));
            NAB_CUDA(cudaMemcpy2DAsync(da.p, ld * 8, a, (m == 1 ? n : (size_t)rsa) * 8, n * 8, m, cudaMemcpyHostToDevice, s));
        } else {
            ha.resize(m * n);
            for (size_t j = 0; j < n; ++j) for (size_t i = 0; i < m; ++i) ha[i + j * m] = a[(ptrdiff_t)i * rsa + (ptrdiff_t)j * csa];
            NAB_TRY(da.alloc(ld * n * 8, s));
            NAB_CUDA(cudaMemcpy2DAsync(da.p, ld * 8, ha.data(), m * 8, m * 8, n, cudaMemcpyHostToDevice, s));
        }
        NAB_CUDA(cudaMemcpyAsync(dx.p, hx.data(), n * 8, cudaMemcpyHostToDevice, s));
    }
    if (beta != 0.0) NAB_CUDA(cudaMemcpyAsync(dy.p, hy.data(), m * 8, cudaMemcpyHostToDevice, s));
    NAB_TRY(gemv_device(s, m, n, alpha, da.as<double>(), drs, dcs, dx.as<double>(), 1, beta, dy.as<double>(), 1));
    NAB_CUDA(cudaMemcpyAsync(hy.data(), dy.p, m * 8, cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < m; ++i) y[(ptrdiff_t)i * incy] = hy[i];
    return NA_OK;
}

}  // extern "C"
