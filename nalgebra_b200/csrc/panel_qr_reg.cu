// panel_qr_reg.cu -- the GEQR2 leaf with its rows in REGISTERS.
//
// Same algorithm, same exchange words and the same results as geqr2_coop_kernel (panel_qr.cu; reference semantics:
// QR::new -> householder::clear_column_unchecked -> reflection_axis_mut -> Reflection::reflect_with_sign,
// /root/reference/src/linalg/qr.rs:55-76, householder.rs:19-85, geometry/reflection.rs:70-83), but a thread owns two whole
// rows of the 32-column panel (128 of its registers) instead of the CTA keeping 672 rows in shared memory.  Applying a
// reflector and accumulating the next column's sums is then pure register arithmetic -- the shared-memory pass was
// 5600 of ~14000 cycles per column.
// 512 rows per CTA: a 65536-row panel takes 128 CTAs, so this leaf is for the drivers that give the panel the whole GPU.
#include <algorithm>
#include <type_traits>
#include <utility>

#include "common.cuh"
#include "kernels.cuh"
#include "panel_qr.cuh"

namespace nab {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F& f) {
    if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

namespace qrr {
constexpr int T = 256;        // threads per CTA
constexpr int RPT = 2;        // rows per thread
constexpr int ROWS = T * RPT; // rows per CTA
constexpr int W = 32;         // panel width the registers are laid out for
}  // namespace qrr

#ifdef NAB_GEQR2_PROF   // per-phase cycle counters of CTA 1 / thread 0 (tools/qr_timing.py), off in the product build
__device__ long long g_geqr2r_prof[16];
#define RPROF(i) do { if (threadIdx.x == 0 && blockIdx.x == 1) { const long long t_ = clock64(); g_geqr2r_prof[i] += t_ - rt_prev; rt_prev = t_; } } while (0)
#else
#define RPROF(i) do { } while (0)
#endif

__global__ void __launch_bounds__(qrr::T, 1) geqr2_reg_kernel(const Geqr2Params p) {
    using namespace qrr;
#ifdef NAB_GEQR2_PROF
    long long rt_prev = clock64();
#endif
    __shared__ double wred[8 * 32];          // per-warp partial sums / reducer scratch
    __shared__ double tot[32];               // totals T_j of the column being received
    __shared__ double rowv[32];              // row c
    __shared__ double dots[32];              // tau * (v^T a_j)
    extern __shared__ double stage[];        // [G][32]: one-stage exchange only (G < 56)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int w = p.w;
    const int r_begin = cta * ROWS;
    const int ncol = min(w, p.m);
    int gr[RPT];
    double a[RPT][W];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        gr[k] = r_begin + k * T + tid;
#pragma unroll
        for (int j = 0; j < W; ++j) a[k][j] = (gr[k] < p.m && j < w) ? p.a[(long long)gr[k] + (long long)j * p.lda] : 0.0;
    }

    // receive column c: totals T_j (identical in every CTA: summed in a fixed order) and row c
    auto receive = [&](int c) {
        const double seq = (double)(p.seq0 + c + 1);
        const int par = c & 1;
        const int np = w - c;                                   // slots c .. w-1
        if (G >= 56) {
            // two-stage exchange: slot j = c + q is summed by ONE reducer CTA (the q-th from the end) and published as
            // a total; every CTA then reads np totals instead of G * np partials
            for (int q = G - 1 - cta; q < np; q += G) {
                const int j = c + q;
                double v = 0.0;
                if (tid < G) {
                    const double2* src = p.xch + ((size_t)par * G + tid) * 32 + j;
                    double y;
                    do { ld_pair_raw(src, v, y); } while (y != seq);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) wred[warp] = v;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    const int nwu = (G + 31) / 32;
                    for (int i = 0; i < nwu; ++i) t += wred[i];
                    st_pair(p.totx + par * 32 + j, t, seq);
                }
                __syncthreads();
            }
            if (tid < np) {
                double x, y;
                const double2* src = p.totx + par * 32 + c + tid;
                do { ld_pair_raw(src, x, y); } while (y != seq);
                tot[c + tid] = x;
            } else if (tid >= 32 && tid < 32 + np) {
                double x, y;
                const double2* src = p.rowc + par * 32 + c + (tid - 32);
                do { ld_pair_raw(src, x, y); } while (y != seq);
                rowv[c + (tid - 32)] = x;
            }
            __syncthreads();
            return;
        }
        const int total = G * np;
        for (int base = 0; base < total; base += 8 * T) {
            double xv[8], yv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * T + tid;
                xv[u] = 0.0; yv[u] = seq;
                if (idx < total) {
                    const int g = idx / np, j = c + (idx - g * np);
                    ld_pair_raw(p.xch + ((size_t)par * G + g) * 32 + j, xv[u], yv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * T + tid;
                if (idx < total) {
                    const int g = idx / np, j = c + (idx - g * np);
                    const double2* src = p.xch + ((size_t)par * G + g) * 32 + j;
                    while (yv[u] != seq) ld_pair_raw(src, xv[u], yv[u]);
                    stage[g * 32 + j] = xv[u];
                }
            }
        }
        if (tid < np) {
            double x, y;
            const double2* src = p.rowc + par * 32 + c + tid;
            do { ld_pair_raw(src, x, y); } while (y != seq);
            rowv[c + tid] = x;
        }
        __syncthreads();
        if (tid < np) {
            double t = 0.0;
            for (int g = 0; g < G; ++g) t += stage[g * 32 + c + tid];
            tot[c + tid] = t;
        }
        __syncthreads();
    };

    // Applies reflector c (c < 0: none) to this thread's rows and accumulates, for the next column cn = c + 1, the sums
    // vals[j] = sum_{r > cn} x[r] * a[r, j] (x = updated column cn); reduces and publishes them with row cn.
    // The columns are handled in four groups of eight: CG = c / 8 is a compile-time constant (columns left of the group
    // are dead code), the position inside the group is a run-time predicate / select.  A fully unrolled column loop
    // (32 copies) made every register index static but was 55 000 instructions -- instruction-fetch bound.
    auto pass = [&](auto cg_tag, int c, double scale) {
        constexpr int J0 = 8 * decltype(cg_tag)::value;
        const int cn = c + 1;                               // J0 <= cn <= J0 + 8
        double vals[W];
#pragma unroll
        for (int j = 0; j < W; ++j) vals[j] = 0.0;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int g = gr[k];
            if (g < p.m && (g >= cn || g == c)) {            // rows above the pivot row are finished (R)
                if (c >= 0) {
                    double ac = 0.0;
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) if (J0 + jj == c) ac = a[k][J0 + jj];
                    const double v = g == c ? 1.0 : ac * scale;     // v[c] = 1; the diagonal slot holds beta (set by the caller)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) if (J0 + jj == c && g != c) a[k][J0 + jj] = v;
                    // dots[j] = 0 for j <= c and j >= w, so the finished columns of the group need no predicate -- unless v
                    // is not finite (0 * inf): then only the columns right of c may change, as in the reference
                    if (isfinite(v)) {
#pragma unroll
                        for (int j = J0; j < W; ++j) a[k][j] = fma(-dots[j], v, a[k][j]);
                    } else {
#pragma unroll
                        for (int j = J0; j < W; ++j) if (j >= cn) a[k][j] = fma(-dots[j], v, a[k][j]);
                    }
                }
                double xs = 0.0;
#pragma unroll
                for (int jj = 0; jj <= 8; ++jj) if (J0 + jj < W && J0 + jj == cn) xs = a[k][J0 + jj < W ? J0 + jj : 0];
                const double x = g > cn ? xs : 0.0;
                // sums of the finished columns (j < cn) are garbage that is never published
#pragma unroll
                for (int j = J0; j < W; ++j) vals[j] = fma(x, a[k][j], vals[j]);
            }
        }
        RPROF(3);
        if (cn >= ncol) return;                              // nothing left to exchange (uniform)
        const double mine = warp_reduce_scatter32(vals, lane);
        wred[warp * 32 + lane] = mine;
        RPROF(4);
        __syncthreads();
        RPROF(5);
        const double seq = (double)(p.seq0 + cn + 1);
        const int par = cn & 1;
        if (tid < 32) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += wred[i * 32 + tid];
            if (tid >= cn && tid < w) st_pair(p.xch + ((size_t)par * G + cta) * 32 + tid, t, seq);
        }
        if (cta == 0 && tid == cn) {                         // the owner of row cn (rows 0..31 are the first rows of CTA 0's threads 0..31)
#pragma unroll
            for (int j = J0; j < W; ++j)
                if (j >= cn && j < w) st_pair(p.rowc + par * 32 + j, a[0][j], seq);
        }
    };

    if (tid < 32) dots[tid] = 0.0;
    __syncthreads();
    pass(std::integral_constant<int, 0>{}, -1, 0.0);         // partial sums of column 0
    __syncthreads();                                         // wred is reused by the reducers in receive()

    auto group = [&](auto cg_tag) {
        constexpr int J0 = 8 * decltype(cg_tag)::value;
#pragma unroll 1
        for (int c = J0; c < J0 + 8 && c < ncol; ++c) {
            RPROF(0);
            receive(c);
            RPROF(1);
            const double alpha = rowv[c], sigma = tot[c];
            const double nrm = sqrt(alpha * alpha + sigma);
            double beta = 0.0, tau = 0.0, scale = 0.0;
            if (nrm != 0.0) {                         // householder.rs:36: only an all-zero column is skipped
                beta = (alpha >= 0.0) ? -nrm : nrm;   // -sign(alpha)*|x|, sign(0) = +1 like simba's to_exp
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (cta == 0 && tid == 0) p.tau[c] = tau;
            __syncthreads();                          // everyone has read rowv/tot before dots is rewritten
            if (tid < 32) dots[tid] = (tid < w && tid > c) ? tau * (rowv[tid] + scale * tot[tid]) : 0.0;   // tau * v^T a_j
            if (cta == 0 && tid == c) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) if (J0 + jj == c) a[0][J0 + jj] = beta;
            }
            __syncthreads();
            RPROF(2);
            pass(cg_tag, c, scale);
            RPROF(6);
            __syncthreads();
            RPROF(7);
        }
    };
    static_for<0, W / 8>(group);

#pragma unroll
    for (int k = 0; k < RPT; ++k)
#pragma unroll
        for (int j = 0; j < W; ++j)
            if (gr[k] < p.m && j < w) p.a[(long long)gr[k] + (long long)j * p.lda] = a[k][j];
}

// CTAs the register-resident leaf needs for an m-row panel; 0 = does not fit the SMs (or the exchange workspace)
int geqr2_reg_grid(size_t m) {
    const size_t G = ceil_div(m, (size_t)qrr::ROWS);
    return G <= (size_t)std::min<int>(ctx().sm_count, (int)kGeqr2MaxCtas) ? (int)G : 0;
}

// Same contract as geqr2_panel (panel_qr.cu): same workspace, same sequence numbers, so leaves of both kinds may
// alternate inside one factorization.
int geqr2_panel_reg(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, double* tau, void* ws, int* seq_state) {
    if (m == 0 || w == 0) return NA_OK;
    if (w > (size_t)qrr::W) { set_error("geqr2 (reg): panel too wide"); return NA_EINVAL; }
    const int G = geqr2_reg_grid(m);
    if (G == 0) { set_error("geqr2 (reg): %zu rows do not fit", m); return NA_EINVAL; }
    Geqr2Params p;
    p.a = a_panel; p.lda = (long long)lda; p.m = (int)m; p.w = (int)w; p.rp = qrr::ROWS; p.tau = tau;
    p.xch = static_cast<double2*>(ws);
    p.rowc = p.xch + 2 * kGeqr2MaxCtas * 32;
    p.totx = p.rowc + 2 * 32;
    p.seq0 = *seq_state;
    *seq_state += (int)w + 2 + ((w & 1) ? 1 : 0);
    const size_t smem = (size_t)(G < 56 ? G : 0) * 32 * sizeof(double);
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)geqr2_reg_kernel, dim3((unsigned)G), dim3(qrr::T), args, smem, st));
    count_launch();
    return NA_OK;
}

}  // namespace nab

#if defined(NAB_GEQR2_PROF) && defined(NAB_DEBUG_HOOKS)   // debug build only (nalgebra_b200/build.py)
extern "C" __attribute__((visibility("default"))) int na_debug_geqr2r_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, nab::g_geqr2r_prof, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(nab::g_geqr2r_prof, z, sizeof(z)); }
    return 0;
}
#endif
