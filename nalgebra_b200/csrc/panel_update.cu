// panel_update.cu -- rank-K update C -= A * B of a tall C for the leaves of the LU panel (K = leaf width: 16, 32, 64).
//
// Reference semantics: the trailing-column updates of gauss_step(_swap) (/root/reference/src/linalg/lu.rs:337-389) for
// K consecutive pivots at once, restricted to the columns of the current outer panel: A22 -= A21 * U12.  With K <= 64
// this is a bandwidth problem (C is read and written once, K FMAs per element); see dmma_stream.cuh.
#include <algorithm>

#include "common.cuh"
#include "dmma_stream.cuh"
#include "kernels.cuh"

namespace nab {

namespace ru {
constexpr int T = 256;
constexpr int NCB = 224;          // columns of C per CTA
constexpr int NP = 228;           // shared-memory row stride of B: = 4 (mod 16)
}  // namespace ru

template <int KS>
__global__ void __launch_bounds__(ru::T, 1) rank_update_kernel(double* __restrict__ c, long long ldc, const double* __restrict__ a, long long lda,
                                                               const double* __restrict__ b, long long ldb, int m, int n, int rows_cta) {
    using namespace ru;
    constexpr int K = 4 * KS;
    extern __shared__ __align__(16) double bs[];            // [K][NP] (+ slack: dmma_stream_update reads up to column 263 of the last row)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.y * NCB, ncb = min(NCB, n - col0);
    // B (K x ncb) -> shared memory, eight independent loads per thread in flight
    for (int i0 = tid; i0 < K * ncb; i0 += 8 * T) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = i0 + u * T;
            t[u] = idx < K * ncb ? b[idx % K + (long long)(col0 + idx / K) * ldb] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = i0 + u * T;
            if (idx < K * ncb) bs[(idx % K) * NP + idx / K] = t[u];
        }
    }
    __syncthreads();
    const int r_cta = blockIdx.x * rows_cta;
    const int nrows = max(0, min(rows_cta, m - r_cta));
    const int g8 = lane >> 2, q4 = lane & 3;
    const double* arow = a + r_cta + g8 + (long long)q4 * lda;
    auto load_a = [&](double (&na)[KS], int r) {
        const bool rok = r + g8 < nrows;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) na[ks] = rok ? -arow[r + (long long)(4 * ks) * lda] : 0.0;
    };
    dmma_stream_update<KS, NP, false, 8>(c + r_cta + (long long)col0 * ldc, ldc, nrows, ncb, bs, lane, warp, T / 32, load_a);
}

// C (m x n, ldc) -= A (m x k, lda) * B (k x n, ldb), all column-major.  Returns NA_OK when it ran, 1 when the shape is
// not one it handles (k not in {16, 32, 64}): the caller then uses the GEMM engine.  max_ctas: SMs it may occupy (0 = all).
int rank_update_small_k(cudaStream_t st, size_t m, size_t k, size_t n, const double* a, size_t lda, const double* b, size_t ldb, double* c,
                        size_t ldc, int max_ctas) {
    using namespace ru;
    if (m == 0 || n == 0) return NA_OK;
    if (k != 16 && k != 32 && k != 64) return 1;
    int budget = ctx().sm_count;
    if (max_ctas > 0) budget = std::min(budget, max_ctas);
    const size_t ncb = ceil_div(n, (size_t)NCB);
    const size_t slabs = std::max<size_t>(1, (size_t)budget / ncb);
    const size_t rows_cta = std::max<size_t>(64, round_up(ceil_div(m, slabs), 64));
    const dim3 grid((unsigned)ceil_div(m, rows_cta), (unsigned)ncb);
    const size_t smem = (k * NP + 64) * sizeof(double);
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(rank_update_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((16 * NP + 64) * sizeof(double)));
        cudaFuncSetAttribute(rank_update_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((32 * NP + 64) * sizeof(double)));
        cudaFuncSetAttribute(rank_update_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((64 * NP + 64) * sizeof(double)));
    });
    if (k == 16) rank_update_kernel<4><<<grid, T, smem, st>>>(c, (long long)ldc, a, (long long)lda, b, (long long)ldb, (int)m, (int)n, (int)rows_cta);
    else if (k == 32) rank_update_kernel<8><<<grid, T, smem, st>>>(c, (long long)ldc, a, (long long)lda, b, (long long)ldb, (int)m, (int)n, (int)rows_cta);
    else rank_update_kernel<16><<<grid, T, smem, st>>>(c, (long long)ldc, a, (long long)lda, b, (long long)ldb, (int)m, (int)n, (int)rows_cta);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab
