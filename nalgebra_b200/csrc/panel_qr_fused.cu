// panel_qr_fused.cu -- the in-panel block reflector of the blocked Householder QR as ONE cooperative kernel.
//
// After a 32-column GEQR2 leaf (classical v / tau form, panel_qr.cu) the rest of the 256-column outer panel is
// updated with  C <- (I - V T^T V^T) C,  C = ml x nc (nc <= 224).  As separate launches this was: clean copy of V +
// Gram partials, Gram finish (S = T^-1 = triu(V^T V, 1) + diag(1/tau)), a split-K GEMM for W = V^T C + its reduction,
// TRTRI + GEMM + copy for X = S^-T W, and a GEMM for C -= V X: nine launches, ~350 us per leaf at ml = 65536, 112
// leaves per 65536 x 4096 factorization -- latency, not bandwidth (one pass over C is 117 MB, ~25 us of HBM time).
// Reference semantics: the trailing-column reflections of QR::new (/root/reference/src/linalg/qr.rs:55-76 ->
// householder.rs:61-85 -> reflection.rs:70-83), 32 reflectors at a time.
//
// Here every CTA owns a slab of rows, every warp a slice of the slab, and
//   1. accumulates its partial  [W | G]_g = V_g^T [C_g | V_g]  (32 x (nc + 32)) on the FP64 tensor pipe, DMMA fragments
//      loaded straight from global memory (V cleaned on the fly: zeros above the diagonal, unit diagonal, zero column
//      when tau = 0); the eight warps' tiles are added in warp order through shared memory,
//   2. publishes it; CTA g then sums slice g of all partials in a fixed order and publishes the totals -- two
//      grid-wide hand-offs through flag words, no atomics,
//   3. builds S from G and tau and solves S^T X = W for its copy of W (32 forward-substitution steps, one thread per
//      column of C),
//   4. streams its slab of C once more:  C_g -= V_g X  (DMMA again: V and C fragments from global memory, X from shared
//      memory, the next 64 columns' loads in flight while the current ones are in the pipe).
// All CTAs must be co-resident (the hand-offs spin): cooperative launch, grid <= the SMs the caller keeps free.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "dmma_stream.cuh"
#include "kernels.cuh"

namespace nab {

namespace lf {
// threads: NW warps (template parameter of the kernel: 8 with explicit double buffering, or 16 with half the registers and
// the warps covering each other's memory stalls), interleaved over the CTA's rows
constexpr int W = 32;             // reflectors per leaf
constexpr int NCX = 256;          // row stride of the [W | G] partials / totals in global memory: nc <= 224, + 32
constexpr int NCXP = 260;         // row stride of X in shared memory: = 4 (mod 16) doubles, so the DMMA B fragments
                                  // (4 consecutive reflectors x 8 consecutive columns) hit 16 distinct 8-byte banks
constexpr int RLD = 36;           // row stride of a warp's 32 x 32 partial tile in shared memory
}  // namespace lf

struct LarfbParams {
    double* a; long long lda;     // leaf origin A[jl, jl]: columns [0, w) hold the reflector vectors, [w, w + nc) is C
    int ml, w, nc, rows_cta;      // rows_cta: multiple of 64 (eight warps x 8-row DMMA tiles)
    const double* tau;            // [w]
    double* part;                 // [G][W * NCX] partial sums
    double* total;                // [W * NCX]
    int* flags;                   // [2][G] hand-off words
    int seq;                      // value the flags take in this launch (monotonic across launches)
};

__device__ __forceinline__ void flag_set(int* f, int v) {
    __threadfence();
    asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(f), "r"(v) : "memory");
}
__device__ __forceinline__ void flag_wait(const int* f, int v) {
    int x;
    do { asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(x) : "l"(f) : "memory"); } while (x < v);
    __threadfence();
}

// Both products run on the FP64 tensor pipe (mma.sync m8n8k4, ptx::dmma884) with fragments loaded straight from global
// memory: a fragment is 4 consecutive rows x 8 columns (or 8 rows x 4 columns), i.e. whole 32-byte sectors, so no
// shared-memory staging and no barrier inside the streaming loops.
#ifdef NAB_LARFB_PROF
__device__ long long g_larfb_prof[16];
#define LF_T(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 1) { long long t_ = clock64(); atomicAdd((unsigned long long*)&g_larfb_prof[i], (unsigned long long)(t_ - lf_t0)); lf_t0 = t_; } } while (0)
#else
#define LF_T(i) do { } while (0)
#endif

template <int NW>
__global__ void __launch_bounds__(32 * NW, 1) larfb_leaf_fused_kernel(const LarfbParams p) {
    using namespace lf;
    constexpr int T = 32 * NW;
    constexpr int PB = NW == 8 ? 4 : 2;              // DMMA k-steps (4 rows each) per phase-1 batch
    constexpr bool PREFETCH = NW == 8;               // explicit double buffering only where the registers allow it
    constexpr int NSUB = T / 64;                     // partial sums per entry in the cross-CTA reduction
#ifdef NAB_LARFB_PROF
    long long lf_t0 = clock64();
#endif
    extern __shared__ __align__(16) double sm[];
    double* red = sm;                                // phase 1: [NW warps][W][RLD] partial tiles
    double* Ws = sm;                                 // later:   [W][NCXP] totals, then X
    double* Ss = sm + W * NCXP;                      //          [W][W + 1]: S(i, j), i <= j
    __shared__ double tau_s[W];
    __shared__ double red4[NSUB][64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, q4 = lane & 3;         // DMMA fragment coordinates
    const int G = gridDim.x, cta = blockIdx.x;
    const int w = p.w, nc = p.nc, ncx = nc + W;
    const long long lda = p.lda;
    const int r_cta = cta * p.rows_cta;                              // the CTA's slab of rows
    const int slab_n = max(0, min(p.rows_cta, p.ml - r_cta));        // valid rows in it
    if (tid < W) tau_s[tid] = tid < w ? p.tau[tid] : 0.0;
    __syncthreads();
    const bool top = cta == 0;                       // the slab holds the triangle above the reflectors' unit heads
    const double* vbase = p.a;
    double* cbase = p.a + (long long)w * lda;

    // ---- phase 1: partial [W | G] = V^T [C | V] over this CTA's rows, one 32-column group of [C | V] at a time.  The
    // slab is cut into 16-row batches (four DMMA k-steps), warp w taking batches w, w + 8, ...: the CTA reads 128
    // consecutive rows of the same columns at a time.  A batch (16 V + 16 C fragment loads per thread) is in flight while
    // the previous one is in the tensor pipe; the first batch of the next group is requested before the cross-warp
    // reduction of the current one.
    {
        bool nz[4];                                  // reflector 8 mt + g8 is present (tau != 0)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) nz[mt] = tau_s[8 * mt + g8] != 0.0;
        const int n_cg = (nc + 31) >> 5;             // groups of C columns; group n_cg is V itself
        double* mypart = p.part + (size_t)cta * (W * NCX);
        const double* vq = vbase + r_cta + q4 + (long long)g8 * lda;
        const double* cq = cbase + r_cta + q4 + (long long)g8 * lda;
        constexpr int BR = 4 * PB;                   // rows per batch
        const int r_first = BR * warp;
        double cav[PB][4], cbv[PB][4], nav[PB][4], nbv[PB][4];
        auto load_batch = [&](double (&av)[PB][4], double (&bv)[PB][4], int grp, int r0) {
            const bool vgrp = grp == n_cg;
#pragma unroll
            for (int s4 = 0; s4 < PB; ++s4) {
                const int r = r0 + 4 * s4;
                const bool rok = r + q4 < slab_n;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) av[s4][mt] = (rok && nz[mt]) ? vq[r + (long long)(8 * mt) * lda] : 0.0;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    bv[s4][nt] = (rok && !vgrp && 32 * grp + 8 * nt + g8 < nc) ? __ldcg(cq + r + (long long)(32 * grp + 8 * nt) * lda) : 0.0;
            }
        };
        if constexpr (PREFETCH) load_batch(cav, cbv, 0, r_first);
        for (int grp = 0; grp <= n_cg; ++grp) {
            double acc[4][4][2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
            const bool vgrp = grp == n_cg;
            for (int r0 = r_first; r0 < slab_n; r0 += BR * NW) {
                if constexpr (PREFETCH) {
                    if (r0 + BR * NW < slab_n) load_batch(nav, nbv, grp, r0 + BR * NW);
                    else if (!vgrp) load_batch(nav, nbv, grp + 1, r_first);
                } else {
                    load_batch(cav, cbv, grp, r0);
                }
#pragma unroll
                for (int s4 = 0; s4 < PB; ++s4) {
                    if (top && r0 < W) {
                        const int row = r0 + 4 * s4 + q4;
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) {
                            const int col = 8 * mt + g8;
                            if (row < col) cav[s4][mt] = 0.0; else if (row == col) cav[s4][mt] = nz[mt] ? 1.0 : 0.0;
                        }
                    }
                    if (vgrp) {
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) cbv[s4][nt] = cav[s4][nt];
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) ptx::dmma884(acc[mt][nt][0], acc[mt][nt][1], cav[s4][mt], cbv[s4][nt]);
                }
                if constexpr (PREFETCH) {
#pragma unroll
                    for (int s4 = 0; s4 < PB; ++s4)
#pragma unroll
                        for (int t = 0; t < 4; ++t) { cav[s4][t] = nav[s4][t]; cbv[s4][t] = nbv[s4][t]; }
                }
            }
            // the eight warps' tiles are added in warp order
            double* mine = red + warp * (W * RLD);
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    *reinterpret_cast<double2*>(mine + (8 * mt + g8) * RLD + 8 * nt + 2 * q4) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            __syncthreads();
            const int cbeg = vgrp ? nc : 32 * grp;
#pragma unroll
            for (int u = 0; u < W * 32 / T; ++u) {
                const int e = tid + u * T, i = e >> 5, c = e & 31;
                double sum = 0.0;
#pragma unroll
                for (int wq = 0; wq < NW; ++wq) sum += red[wq * (W * RLD) + i * RLD + c];
                if (cbeg + c < (vgrp ? ncx : nc)) mypart[i * NCX + cbeg + c] = sum;
            }
            __syncthreads();
        }
    }
    LF_T(0);
    // ---- hand-off 1: publish the partial, then CTA g sums slice g of all partials in CTA order (four quarter sums
    // of consecutive CTAs, added in order: a fixed summation shape for a given grid)
    if (tid == 0) flag_set(p.flags + cta, p.seq);
    {
        const int tot = W * ncx;
        const int E = (tot + G - 1) / G;                         // entries per reducer CTA (<= 64: G >= 128 or ncx small)
        const int e0 = cta * E, e1 = min(tot, e0 + E);
        if (e0 < e1) {
            if (tid < G) flag_wait(p.flags + tid, p.seq);
            __syncthreads();
            for (int eb = e0; eb < e1; eb += 64) {
                const int sub = tid >> 6, e = eb + (tid & 63);
                double s = 0.0;
                if (e < e1) {
                    const int i = e / ncx, c = e - i * ncx;
                    const double* src = p.part + i * NCX + c;
                    const int ga = G * sub / NSUB, gb = G * (sub + 1) / NSUB;
                    int g = ga;
                    for (; g + 10 <= gb; g += 10) {
                        double t[10];
#pragma unroll
                        for (int u = 0; u < 10; ++u) t[u] = __ldcg(src + (size_t)(g + u) * (W * NCX));
#pragma unroll
                        for (int u = 0; u < 10; ++u) s += t[u];
                    }
                    for (; g < gb; ++g) s += __ldcg(src + (size_t)g * (W * NCX));
                }
                red4[sub][tid & 63] = s;
                __syncthreads();
                if (tid < 64 && e < e1) {
                    const int i = e / ncx, c = e - i * ncx;
                    double t = 0.0;
#pragma unroll
                    for (int q = 0; q < NSUB; ++q) t += red4[q][tid];
                    p.total[i * NCX + c] = t;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (tid == 0) flag_set(p.flags + G + cta, p.seq);
    }
    LF_T(1);
    // ---- hand-off 2: everybody takes the totals once every reducer is done
    if (tid < G) flag_wait(p.flags + G + tid, p.seq);
    __syncthreads();
    for (int idx = tid; idx < W * ncx; idx += T) {
        const int i = idx / ncx, c = idx - i * ncx;
        Ws[i * NCXP + c] = __ldcg(p.total + i * NCX + c);
    }
    __syncthreads();
    LF_T(2);
    // ---- phase 3: S = triu(G, 1) + diag(1 / tau) (tau = 0 -> 1: that column of V is zero); solve S^T X = W
    for (int idx = tid; idx < W * W; idx += T) {
        const int i = idx >> 5, j = idx & 31;
        double v = 0.0;
        if (i < j) v = Ws[i * NCXP + nc + j];
        else if (i == j) v = tau_s[i] != 0.0 ? 1.0 / tau_s[i] : 1.0;
        Ss[i * (W + 1) + j] = v;
    }
    __syncthreads();
    if (tid < nc) {          // one thread per column of C: x_i = (w_i - sum_{j<i} S(j, i) x_j) / S(i, i)
        double xv[W];        // (with 16 warps this array spills to local memory: 400 bytes once per launch, cheaper than
                             // walking the column in shared memory -- 10 000 vs 20 000 cycles)
#pragma unroll
        for (int i = 0; i < W; ++i) xv[i] = Ws[i * NCXP + tid];
#pragma unroll
        for (int i = 0; i < W; ++i) {
            double s = xv[i];
#pragma unroll
            for (int j = 0; j < i; ++j) s = fma(-Ss[j * (W + 1) + i], xv[j], s);
            xv[i] = s / Ss[i * (W + 1) + i];
        }
#pragma unroll
        for (int i = 0; i < W; ++i) Ws[i * NCXP + tid] = xv[i];
    }
    __syncthreads();
    LF_T(3);
    // ---- phase 4: C -= V X on the warp's row slice (dmma_stream.cuh), last rows first: the most recently read lines
    // are the likeliest L2 hits
    {
        bool nzk[8];                                 // reflector 4 ks + q4 present
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) nzk[ks] = tau_s[4 * ks + q4] != 0.0;
        const double* vrow = vbase + r_cta + g8 + (long long)q4 * lda;
        auto load_a = [&](double (&na)[8], int r) {
            const bool rok = r + g8 < slab_n;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) na[ks] = (rok && nzk[ks]) ? -vrow[r + (long long)(4 * ks) * lda] : 0.0;
            if (top && r < W) {
                const int row = r + g8;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const int col = 4 * ks + q4;
                    if (row < col) na[ks] = 0.0; else if (row == col) na[ks] = nzk[ks] ? -1.0 : 0.0;
                }
            }
        };
        dmma_stream_update<8, NCXP, true, (NW == 8 ? 8 : 4)>(cbase + r_cta, lda, slab_n, nc, Ws, lane, warp, NW, load_a);
    }
    LF_T(4);
}

constexpr size_t kLarfbMaxCtas = 160;
size_t larfb_fused_workspace_bytes() {
    return (kLarfbMaxCtas + 1) * (size_t)lf::W * lf::NCX * sizeof(double) + 2 * kLarfbMaxCtas * sizeof(int) + 256;
}

// C <- (I - V T^T V^T) C for the leaf at `a_leaf` (ml x (w + nc): reflector vectors, then C).  ws: zero-initialised
// workspace of larfb_fused_workspace_bytes(); *seq (host) counts the launches that used it.  max_ctas: SMs the caller
// keeps free for this cooperative launch (0 = all).  Returns NA_EINVAL when the shape does not fit (nc > 224, w > 32).
int larfb_leaf_fused(cudaStream_t st, double* a_leaf, size_t lda, size_t ml, size_t w, size_t nc, const double* tau, void* ws, int* seq,
                     int max_ctas) {
    using namespace lf;
    if (ml == 0 || w == 0 || nc == 0) return NA_OK;
    if (w > (size_t)W || nc > (size_t)(NCX - W)) { set_error("larfb_fused: %zu reflectors x %zu columns do not fit", w, nc); return NA_EINVAL; }
    int G = ctx().sm_count;
    if (max_ctas > 0) G = std::min(G, max_ctas);
    G = (int)std::min<size_t>((size_t)G, std::min<size_t>(kLarfbMaxCtas, ceil_div(ml, (size_t)64)));
    const size_t rows_cta = round_up(ceil_div(ml, (size_t)G), (size_t)64);
    G = (int)ceil_div(ml, rows_cta);
    static const int nw = [] { const char* e = getenv("NAB_LARFB_WARPS"); return (e && atoi(e) == 8) ? 8 : 16; }();
    const size_t smem = std::max<size_t>((size_t)nw * W * RLD, (size_t)W * NCXP + W * (W + 1)) * sizeof(double);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(larfb_leaf_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(std::max<size_t>(8 * (size_t)W * RLD, (size_t)W * NCXP + W * (W + 1)) * sizeof(double)));
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(larfb_leaf_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(std::max<size_t>(16 * (size_t)W * RLD, (size_t)W * NCXP + W * (W + 1)) * sizeof(double)));
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(larfb_fused)", __FILE__, __LINE__);
    LarfbParams p;
    p.a = a_leaf; p.lda = (long long)lda; p.ml = (int)ml; p.w = (int)w; p.nc = (int)nc; p.rows_cta = (int)rows_cta; p.tau = tau;
    p.part = static_cast<double*>(ws);
    p.total = p.part + kLarfbMaxCtas * (size_t)W * NCX;
    p.flags = reinterpret_cast<int*>(p.total + (size_t)W * NCX);
    p.seq = ++*seq;
    void* args[] = {(void*)&p};
    if (nw == 8) NAB_CUDA(cudaLaunchCooperativeKernel((void*)larfb_leaf_fused_kernel<8>, dim3((unsigned)G), dim3(256), args, smem, st));
    else NAB_CUDA(cudaLaunchCooperativeKernel((void*)larfb_leaf_fused_kernel<16>, dim3((unsigned)G), dim3(512), args, smem, st));
    count_launch();
    return NA_OK;
}

}  // namespace nab

#if defined(NAB_LARFB_PROF) && defined(NAB_DEBUG_HOOKS)   // debug build only (nalgebra_b200/build.py)
extern "C" __attribute__((visibility("default"))) int na_debug_larfb_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, nab::g_larfb_prof, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(nab::g_larfb_prof, z, sizeof(z)); }
    return 0;
}
#endif
