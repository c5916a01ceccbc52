// sgemm.cu -- f32 GEMM (replaces matrixmultiply::sgemm, /root/reference/src/base/blas_uninit.rs:276-291).
//
// FFMA register-tiled kernel: CTA tile 128x128x16, 256 threads, 8x8 accumulators per thread,
// operands staged through shared memory with register prefetch of the next k-slab.  Arbitrary
// element strides on A, B and C are handled directly in the loads/stores (the load mapping follows
// whichever stride of the operand is the small one, so NN/NT/TN/TT are all coalesced).
// Results are plain f32 FFMA accumulations (within the reference tolerance by construction; a
// tcgen05 3xTF32 variant is the planned upgrade).
#include <algorithm>
#include <vector>

#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "kernels.cuh"

namespace nab {

namespace scfg {
constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256;
}

struct SgemmParams {
    int M, N, K;
    const float* A; long long rsa, csa;
    const float* B; long long rsb, csb;
    float* C; long long rsc, csc;
    float alpha, beta;
    int a_m_fast, b_n_fast;   // which index is the small-stride one
};

__global__ void __launch_bounds__(scfg::THREADS, 2) sgemm_ffma_kernel(const SgemmParams p) {
    using namespace scfg;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM, n0 = (long long)blockIdx.y * BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto load_slab = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * THREADS;                 // 0..2047
            int mm, kk;
            if (p.a_m_fast) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            const long long gm = m0 + mm; const int gk = k0 + kk;
            ra[i] = (gm < p.M && gk < p.K) ? p.A[gm * p.rsa + gk * p.csa] : 0.f;
            int nn, kb;
            if (p.b_n_fast) { nn = e % BN; kb = e / BN; } else { kb = e % BK; nn = e / BK; }
            const long long gn = n0 + nn; const int gkb = k0 + kb;
            rb[i] = (gn < p.N && gkb < p.K) ? p.B[gkb * p.rsb + gn * p.csb] : 0.f;
        }
    };
    auto store_slab = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * THREADS;
            int mm, kk;
            if (p.a_m_fast) { mm = e % BM; kk = e / BM; } else { kk = e % BK; mm = e / BK; }
            As[kk][mm] = ra[i];
            int nn, kb;
            if (p.b_n_fast) { nn = e % BN; kb = e / BN; } else { kb = e % BK; nn = e / BK; }
            Bs[kb][nn] = rb[i];
        }
    };

    load_slab(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        __syncthreads();
        store_slab();
        __syncthreads();
        if (k0 + BK < p.K) load_slab(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][ty * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    const bool use_beta = p.beta != 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long long gn = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (gn >= p.N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long gm = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (gm >= p.M) continue;
            float* c = p.C + gm * p.rsc + gn * p.csc;
            float v = p.alpha * acc[i][j];
            if (use_beta) v += p.beta * *c;     // C is never read when beta == 0 (blas_uninit.rs:182)
            *c = v;
        }
    }
}

__global__ void sscale_strided_kernel(float* __restrict__ c, long long rs, long long cs, long long rows, long long cols, float beta) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % rows, col = idx / rows;
        float* p = c + r * rs + col * cs;
        *p = (beta == 0.f) ? 0.f : (*p * beta);
    }
}

static long long labs_ll(long long x) { return x < 0 ? -x : x; }

int sgemm_device(cudaStream_t s, size_t m, size_t k, size_t n, float alpha, const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const float* b, ptrdiff_t rsb, ptrdiff_t csb, float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc) {
    NAB_TRY(ensure_init());
    if (m == 0 || n == 0) return NA_OK;
    if (k == 0) {
        const int blocks = (int)std::min<size_t>(ceil_div(m * n, 256), (size_t)ctx().sm_count * 8);
        sscale_strided_kernel<<<blocks, 256, 0, s>>>(c, rsc, csc, (long long)m, (long long)n, beta);
        NAB_LAUNCH_CHECK();
        return NA_OK;
    }
    if (m > 0x7fffff00ull || n > 0x7fffff00ull || k > 0x7fffff00ull) { set_error("sgemm: dimension exceeds 2^31"); return NA_EINVAL; }
    // Tensor-core path (sgemm_tc.cu: TMA + tcgen05.mma kind::tf32, 3xTF32, TMEM accumulators) for everything but tiny
    // products, where one FFMA tile kernel launch beats pack + TMA setup.  NAB_SGEMM=ffma / tc forces a path.
    static const int force = [] { const char* e = getenv("NAB_SGEMM"); return !e ? 0 : (strcmp(e, "ffma") == 0 ? 1 : (strcmp(e, "tc") == 0 ? 2 : 0)); }();
    const bool tiny = (double)m * (double)n * (double)k < 64.0 * 64.0 * 64.0;
    if (force == 2 || (force == 0 && !tiny)) {
        if (rsc == 1 && (csc >= (ptrdiff_t)m || n == 1))
            return sgemm_tc_colmajor(s, m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, (size_t)(n == 1 ? std::max<ptrdiff_t>(csc, (ptrdiff_t)m) : csc));
        if (csc == 1 && (rsc >= (ptrdiff_t)n || m == 1))       // row-major C: C^T = B^T A^T
            return sgemm_tc_colmajor(s, n, k, m, alpha, b, csb, rsb, a, csa, rsa, beta, c, (size_t)(m == 1 ? std::max<ptrdiff_t>(rsc, (ptrdiff_t)n) : rsc));
        // general C strides: fall through to the strided FFMA kernel (rare: views of views as the output)
    }
    SgemmParams p;
    p.M = (int)m; p.N = (int)n; p.K = (int)k;
    p.A = a; p.rsa = rsa; p.csa = csa; p.B = b; p.rsb = rsb; p.csb = csb; p.C = c; p.rsc = rsc; p.csc = csc;
    p.alpha = alpha; p.beta = beta;
    p.a_m_fast = labs_ll(rsa) <= labs_ll(csa) ? 1 : 0;
    p.b_n_fast = labs_ll(csb) <= labs_ll(rsb) ? 1 : 0;
    dim3 grid((unsigned)ceil_div(m, scfg::BM), (unsigned)ceil_div(n, scfg::BN));
    if (grid.y > 65535) { set_error("sgemm: n too large for this kernel"); return NA_EINVAL; }
    sgemm_ffma_kernel<<<grid, scfg::THREADS, 0, s>>>(p);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab

using namespace nab;


extern "C" {

int na_sgemm_dev(size_t m, size_t k, size_t n, float alpha, const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                 const float* b, ptrdiff_t rsb, ptrdiff_t csb, float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc, void* stream) {
    return sgemm_device(static_cast<cudaStream_t>(stream), m, k, n, alpha, a, rsa, csa, b, rsb, csb, beta, c, rsc, csc);
}

int na_sgemm(size_t m, size_t k, size_t n, float alpha, const float* a, ptrdiff_t rsa, ptrdiff_t csa,
             const float* b, ptrdiff_t rsb, ptrdiff_t csb, float beta, float* c, ptrdiff_t rsc, ptrdiff_t csc) {
    NAB_TRY(ensure_init());
    if (m == 0 || n == 0) return NA_OK;
    if ((k && (!a || !b)) || !c) { set_error("sgemm: null pointer"); return NA_EINVAL; }
    if ((m > 1 && rsc == 0) || (n > 1 && csc == 0)) { set_error("sgemm: c has a zero stride"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    // Unit-stride views (column- or row-major: everything nalgebra's VecStorage and its transposed views produce) are
    // staged with one 2-D copy each and keep their strides on the device (the pack pass of the tensor-core path reads
    // any strides); only views with two non-unit strides are gathered on the host.
    struct StagedF {
        Scratch buf; ptrdiff_t rs = 1, cs = 0; std::vector<float> tmp;
    };
    auto stage_in = [&](StagedF& st, const float* h, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols, bool upload) -> int {
        if (rs == 1 && (cs >= (ptrdiff_t)rows || cols == 1)) {
            st.rs = 1; st.cs = (ptrdiff_t)rows;
            NAB_TRY(st.buf.alloc(rows * cols * sizeof(float), s));
            if (upload) NAB_CUDA(cudaMemcpy2DAsync(st.buf.p, rows * 4, h, (cols == 1 ? rows : (size_t)cs) * 4, rows * 4, cols, cudaMemcpyHostToDevice, s));
            return NA_OK;
        }
        if (cs == 1 && (rs >= (ptrdiff_t)cols || rows == 1)) {
            st.rs = (ptrdiff_t)cols; st.cs = 1;
            NAB_TRY(st.buf.alloc(rows * cols * sizeof(float), s));
            if (upload) NAB_CUDA(cudaMemcpy2DAsync(st.buf.p, cols * 4, h, (rows == 1 ? cols : (size_t)rs) * 4, cols * 4, rows, cudaMemcpyHostToDevice, s));
            return NA_OK;
        }
        st.rs = 1; st.cs = (ptrdiff_t)rows;
        NAB_TRY(st.buf.alloc(rows * cols * sizeof(float), s));
        if (upload) {
            st.tmp.resize(rows * cols);
            for (size_t j = 0; j < cols; ++j)
                for (size_t i = 0; i < rows; ++i) st.tmp[i + j * rows] = h[(ptrdiff_t)i * rs + (ptrdiff_t)j * cs];
            NAB_CUDA(cudaMemcpyAsync(st.buf.p, st.tmp.data(), rows * cols * sizeof(float), cudaMemcpyHostToDevice, s));
        }
        return NA_OK;
    };
    StagedF sa, sb, sc;
    if (k) { NAB_TRY(stage_in(sa, a, rsa, csa, m, k, true)); NAB_TRY(stage_in(sb, b, rsb, csb, k, n, true)); }
    NAB_TRY(stage_in(sc, c, rsc, csc, m, n, beta != 0.f));        // C is not read (nor uploaded) when beta == 0
    NAB_TRY(sgemm_device(s, m, k, n, alpha, sa.buf.as<float>(), sa.rs, sa.cs, sb.buf.as<float>(), sb.rs, sb.cs, beta,
                         sc.buf.as<float>(), sc.rs, sc.cs));
    if (sc.rs == 1 && rsc == 1 && (csc >= (ptrdiff_t)m || n == 1)) {
        NAB_CUDA(cudaMemcpy2DAsync(c, (n == 1 ? m : (size_t)csc) * 4, sc.buf.p, m * 4, m * 4, n, cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
    } else if (sc.cs == 1 && csc == 1) {
        NAB_CUDA(cudaMemcpy2DAsync(c, (m == 1 ? n : (size_t)rsc) * 4, sc.buf.p, n * 4, n * 4, m, cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
    } else {
        std::vector<float> hc(m * n);
        NAB_CUDA(cudaMemcpyAsync(hc.data(), sc.buf.p, m * n * sizeof(float), cudaMemcpyDeviceToHost, s));
        NAB_CUDA(cudaStreamSynchronize(s));
        for (size_t j = 0; j < n; ++j)
            for (size_t i = 0; i < m; ++i) c[(ptrdiff_t)i * rsc + (ptrdiff_t)j * csc] = hc[i + j * m];
    }
    return NA_OK;
}

}  // extern "C"
