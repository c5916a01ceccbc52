// panel_qr.cu -- Householder QR panel kernels.
//
// Reference semantics: QR::new -> householder::clear_column_unchecked -> reflection_axis_mut ->
// Reflection::reflect_with_sign (/root/reference/src/linalg/qr.rs:55-76, householder.rs:19-85,
// geometry/reflection.rs:70-83).  The blocked algorithm works in the classical (v, tau, beta)
// convention -- v[0] = 1, H = I - tau v v^T, beta = -sign(alpha)*|x| -- because that is what the
// compact-WY trailing update needs, and converts ONCE at the end to nalgebra's storage
// (unit-2-norm axis u_i = -sign(diag_i) * v_i * sqrt(tau_i/2), diag_i = c_i*beta_i, strict upper row
// i scaled by c_{i+1}; c_{i+1} = sign(beta_i), c_0 = 1).  The mapping is derived in DESIGN.md §3.4
// and checked against the oracle in tests/.  Unlike LAPACK's dlarfg there is no tau = 0 shortcut
// for a column whose sub-diagonal part is zero but whose leading entry is not: nalgebra reflects
// such a column (householder.rs:36-48 only skips when the whole column is zero), so do we.
#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"
#include "panel_qr.cuh"

namespace nab {

// ------------------------------------------------------------------------------------------------
// GEQR2: m x w panel (w <= 32), rows distributed over G co-resident CTAs, resident in shared memory;
// ONE grid-wide exchange per column.
//
// For column c every CTA needs  alpha = a[c,c],  sigma = sum_{r>c} a[r,c]^2,  S_j = sum_{r>c} a[r,c]*a[r,j]
// and the row a[c, j], j > c.  With those it forms the reflector (beta, tau, scale) and the products
// v^T a_j = a[c,j] + scale*S_j locally.  The partial sums for column c+1 are accumulated in the SAME
// pass that applies reflector c (column c+1 is updated first and kept in registers), reduced with a
// butterfly reduce-scatter, and published together with row c+1 by its owner.
//
// The exchange is self-validating: every published double travels with a sequence number in the
// same 16-byte word, so there is no separate barrier, fence or counter: readers poll the words they
// need until the sequence number matches.  Two buffers (by column parity) are enough because a CTA
// cannot publish column c+2 before every CTA has published c+1, i.e. finished reading c.
// ------------------------------------------------------------------------------------------------
// Geqr2Params, st_pair / ld_pair_raw and warp_reduce_scatter32 live in panel_qr.cuh (shared with the register-resident leaf)
__device__ long long g_geqr2_prof[16];
#ifdef NAB_GEQR2_PROF   // per-phase cycle counters of CTA 0 / thread 0 (tools/qr_timing.py), off in the product build
#define QPROF(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); g_geqr2_prof[i] += t_ - qt_prev; qt_prev = t_; } } while (0)
#else
#define QPROF(i) do { (void)qt_prev; } while (0)
#endif

__global__ void __launch_bounds__(256, 1) geqr2_coop_kernel(const Geqr2Params p) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int w = p.w, rp = p.rp;
    const int r_begin = cta * rp;
    const int nrows = max(0, min(rp, p.m - r_begin));
    double* s = sm;                          // [w][rp]
    double* wred = s + (size_t)w * rp;       // [8][32] per-warp partial sums
    double* tot = wred + 8 * 32;             // [32] totals T_j
    double* rowv = tot + 32;                 // [32] row c values
    double* dots = rowv + 32;                // [32] tau * (v^T a_j)
    double* stage = dots + 32;               // [G][32]

    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) s[r + c * rp] = p.a[(long long)(r_begin + r) + (long long)c * p.lda];
    __syncthreads();
    const int ncol = min(w, p.m);
    long long qt_prev = clock64();

    // One pass over the local rows: applies reflector `c` (c < 0: nothing to apply) and accumulates, for
    // the next column cn = c + 1, vals[j] = sum_{r > cn} x[r] * a[r, j] (j >= cn) with x = updated column cn.
    auto pass = [&](int c, double scale, double tau_unused) {
        (void)tau_unused;
        const int cn = c + 1;
        double vals[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) vals[j] = 0.0;
        // (measured: this loop was 8500 of 16500 cycles per column at m = 65536 whatever the width -- every thread
        // issued all 32 predicated column bodies and re-read dots[j] per row; now tau*v^T a_j sits in registers and
        // finished 8-column groups are skipped by a uniform branch)
        double dj[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) dj[j] = (c >= 0 && j >= cn && j < w) ? dots[j] : 0.0;
        for (int r = tid; r < nrows; r += nt) {
            const int gr = r_begin + r;
            if (gr < cn && !(gr == c)) continue;            // finished rows (R part) above the pivot row
            double v = 0.0;
            if (c >= 0) {
                if (gr == c) v = 1.0;                        // v[c] = 1; the diagonal slot is set by the caller
                else { v = s[r + c * rp] * scale; s[r + c * rp] = v; }
            }
            double x = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (8 * b + 7 < cn || 8 * b >= w) continue;  // whole group finished / beyond the panel (uniform)
                // the 8 loads of a group are issued before any store of the group: interleaved, every load waited
                // for the previous store (possible aliasing) and the loop ran at shared-memory latency
                double aj[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = 8 * b + jj;
                    aj[jj] = (j >= cn && j < w) ? s[r + j * rp] : 0.0;
                }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = 8 * b + jj;
                    if (j < cn || j >= w) continue;          // columns right of c only
                    if (c >= 0) aj[jj] -= dj[j] * v;
                    if (j == cn) x = (gr > cn) ? aj[jj] : 0.0;   // rows below the next pivot row feed the sums
                    vals[j] += x * aj[jj];
                }
                if (c >= 0) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = 8 * b + jj;
                        if (j >= cn && j < w) s[r + j * rp] = aj[jj];
                    }
                }
            }
        }
        QPROF(4);
        const double mine = warp_reduce_scatter32(vals, lane);
        wred[warp * 32 + lane] = mine;
        QPROF(5);
        __syncthreads();
        QPROF(6);
        // cross-warp sum and publish: slot j = column j (slot cn = sigma), plus row cn from its owner
        if (cn < ncol) {
            const double seq = (double)(p.seq0 + cn + 1);
            const int par = cn & 1;
            if (tid < 32) {
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) t += wred[i * 32 + tid];
                if (tid >= cn && tid < w) st_pair(p.xch + ((size_t)par * G + cta) * 32 + tid, t, seq);
            } else if (tid < 64) {
                const int j = tid - 32;
                if (cn >= r_begin && cn < r_begin + nrows && j >= cn && j < w) st_pair(p.rowc + par * 32 + j, s[(cn - r_begin) + j * rp], seq);
            }
        }
    };

    // receive column c: totals T_j (identical in every CTA: summed in CTA order) and row c
    auto receive = [&](int c) {
        const double seq = (double)(p.seq0 + c + 1);
        const int par = c & 1;
        const int np = w - c;                                   // slots c .. w-1
        if (G >= 56) {    // below that the one-stage all-to-all is faster (measured: G = 25: 2600 vs 4500 cycles; G = 98: 5900 vs 3600)
            // Two-stage exchange (reduce-scatter + all-gather through L2): slot j = c + q is summed by ONE reducer CTA
            // (the q-th from the end, away from CTA 0 which owns the pivot rows) in a fixed order and published as a
            // total; every CTA then reads np totals instead of G * np partials (98 x 32 pairs = 50 KB per CTA and
            // column at m = 65536: measured 5900 of 16500 cycles per column).
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
        // A thread owns up to G*np/256 slots (12 at G = 98, np = 32).  Their loads are issued in batches of 8 before
        // any sequence number is checked: polled one after the other, every slot cost a full L2 round trip.
        const int total = G * np;
        for (int base = 0; base < total; base += 8 * nt) {
            double xv[8], yv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * nt + tid;
                xv[u] = 0.0; yv[u] = seq;
                if (idx < total) {
                    const int g = idx / np, j = c + (idx - g * np);
                    ld_pair_raw(p.xch + ((size_t)par * G + g) * 32 + j, xv[u], yv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * nt + tid;
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

    pass(-1, 0.0, 0.0);                                          // partial sums of column 0
    __syncthreads();                                             // wred is reused by the reducers in receive()
    for (int c = 0; c < ncol; ++c) {
        QPROF(0);
        receive(c);
        QPROF(1);
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
        if (tid < w && tid > c) dots[tid] = tau * (rowv[tid] + scale * tot[tid]);   // tau * v^T a_j
        if (tid == 0 && c >= r_begin && c < r_begin + nrows) s[(c - r_begin) + c * rp] = beta;
        __syncthreads();
        QPROF(2);
        if (c + 1 < w || true) pass(c, scale, tau);
        QPROF(7);
        __syncthreads();
        QPROF(3);
    }
    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) p.a[(long long)(r_begin + r) + (long long)c * p.lda] = s[r + c * rp];
}

size_t geqr2_workspace_bytes() { return (2 * kGeqr2MaxCtas * 32 + 2 * 32 + 2 * 32) * sizeof(double2); }

// CTAs (and rows per CTA) the cooperative GEQR2 launch uses for an m x w panel; 0 if it does not fit.
static size_t geqr2_grid_rows(size_t m, size_t w, size_t& rp_out) {
    const int sms = ctx().sm_count;
    size_t G = 1, rp = 0;
    for (;; ++G) {
        if (G > (size_t)std::min<int>(sms, (int)kGeqr2MaxCtas)) return 0;
        rp = round_up(ceil_div(m, G), 32);
        const size_t bytes = (w * rp + 352 + G * 32) * sizeof(double);
        if (bytes <= 200 * 1024 && (rp <= 768 || G * 2 > (size_t)sms)) break;   // prefer <= 768 rows per CTA when SMs allow
    }
    rp_out = rp;
    return ceil_div(m, rp);
}
int geqr2_grid(size_t m, size_t w) { size_t rp; return (int)geqr2_grid_rows(m, w, rp); }

// *seq_state (host) carries the sequence numbers consumed so far in this (zero-initialised) workspace.
int geqr2_panel(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, double* tau, void* ws, int* seq_state) {
    if (m == 0 || w == 0) return NA_OK;
    if (w > (size_t)kQrLeaf) { set_error("geqr2: panel too wide"); return NA_EINVAL; }
    // smem: w*rp + 8*32 + 3*32 + G*32 doubles
    size_t rp = 0;
    const size_t G = geqr2_grid_rows(m, w, rp);
    if (G == 0) { set_error("geqr2: %zu x %zu panel does not fit in shared memory", m, w); return NA_EINVAL; }
    const size_t smem = (w * rp + 352 + G * 32) * sizeof(double);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(geqr2_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    Geqr2Params p;
    p.a = a_panel; p.lda = (long long)lda; p.m = (int)m; p.w = (int)w; p.rp = (int)rp; p.tau = tau;
    p.xch = static_cast<double2*>(ws);
    p.rowc = p.xch + 2 * kGeqr2MaxCtas * 32;
    p.totx = p.rowc + 2 * 32;
    p.seq0 = *seq_state;
    *seq_state += (int)w + 2 + ((w & 1) ? 1 : 0);
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)geqr2_coop_kernel, dim3((unsigned)G), dim3(256), args, smem, st));
    count_launch();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// V extraction: vw(i,j) = 1 (i == j), panel(i,j) (i > j), 0 (i < j); a column whose tau is 0 is
// zeroed entirely (H = I).  mode 1 (nalgebra axes): the diagonal entry is kept (u includes its
// first component) and tau is implied by diag: tau_j = diag[j] != 0 ? 2 : 0.
// ------------------------------------------------------------------------------------------------
__global__ void extract_v_kernel(double* __restrict__ vw, long long ldv, const double* __restrict__ a, long long lda,
                                 long long m, int w, const double* __restrict__ tau, int mode) {
    const long long total = m * w;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % m; const int j = (int)(idx / m);
        double v = 0.0;
        if (tau[j] != 0.0) {
            if (i > j) v = a[i + j * lda];
            else if (i == j) v = mode ? a[i + j * lda] : 1.0;
        }
        vw[i + j * ldv] = v;
    }
}
int extract_v(cudaStream_t st, double* vw, size_t ldv, const double* a, size_t lda, size_t m, size_t w, const double* tau, int mode) {
    if (m == 0 || w == 0) return NA_OK;
    extract_v_kernel<<<(int)std::min<size_t>(ceil_div(m * w, 256), (size_t)ctx().sm_count * 16), 256, 0, st>>>(
        vw, (long long)ldv, a, (long long)lda, (long long)m, (int)w, tau, mode);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Leaf panels (w <= 32): V extraction fused with the Gram matrix and S = T^-1 = triu(V^T V, 1) + diag(1/tau).
// Replaces extract_v + a 32 x 32 x m split-K GEMM (a 128 x 128 tile padded 16x, 148 K-slices) + its
// reduction + build_s (160 us per leaf) by two small launches: each CTA cleans chunks of 128 rows into vw
// (16 independent loads per thread in flight), keeps the chunk in shared memory, accumulates its 32 x 32
// partial Gram matrix (thread = one row i and four columns j); gram_finish adds the partials in CTA order
// (deterministic) and writes S.
// ------------------------------------------------------------------------------------------------
constexpr int kGramCtas = 148;
__global__ void __launch_bounds__(256) extract_v_gram_kernel(double* __restrict__ vw, long long ldv, const double* __restrict__ a,
                                                             long long lda, long long m, int w, const double* __restrict__ tau,
                                                             double* __restrict__ part) {
    __shared__ double tile[128 * 33];
    __shared__ double tau_s[32];
    const int tid = threadIdx.x, i = tid & 31, j0 = tid >> 5;
    if (tid < 32) tau_s[tid] = tid < w ? tau[tid] : 0.0;
    __syncthreads();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const long long nchunks = (m + 127) / 128;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long c0 = ch * 128;
        const int nr = (int)min((long long)128, m - c0);
        // 16 independent loads per thread in flight (the chunk is 128 rows x 32 columns), then the stores
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = tid + u * 256, r = idx & 127, j = idx >> 7;
            const long long gr = c0 + r;
            v[u] = (j < w && r < nr && gr > j && tau_s[j] != 0.0) ? a[gr + j * lda] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = tid + u * 256, r = idx & 127, j = idx >> 7;
            const long long gr = c0 + r;
            if (j < w) {
                if (gr == j && tau_s[j] != 0.0) v[u] = 1.0;
                if (r < nr) vw[gr + j * ldv] = v[u];
                tile[r * 33 + j] = v[u];
            }
        }
        __syncthreads();
        if (i < w) {
#pragma unroll 4
            for (int r = 0; r < nr; ++r) {
                const double vi = tile[r * 33 + i];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = fma(vi, tile[r * 33 + j0 + 8 * k], acc[k]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) part[(size_t)blockIdx.x * 1024 + i * 32 + j0 + 8 * k] = acc[k];
}

// S (w x w, ld lds) = triu(sum of the per-CTA partial Gram matrices, 1) + diag(1/tau); CTA order (deterministic).
__global__ void __launch_bounds__(1024) gram_finish_kernel(double* __restrict__ smat, long long lds, int w, const double* __restrict__ tau,
                                                           const double* __restrict__ part, int G) {
    const int i = threadIdx.x & 31, j = threadIdx.x >> 5;
    if (i >= w || j >= w) return;
    double v = 0.0;
    if (i < j) {
        const double* g0 = part + i * 32 + j;
        int g = 0;
        for (; g + 8 <= G; g += 8) {
            double t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = g0[(size_t)(g + k) * 1024];
#pragma unroll
            for (int k = 0; k < 8; ++k) v += t[k];
        }
        for (; g < G; ++g) v += g0[(size_t)g * 1024];
    } else if (i == j) {
        v = (tau[j] != 0.0) ? 1.0 / tau[j] : 1.0;
    }
    smat[i + j * lds] = v;
}
size_t extract_v_gram_workspace_bytes() { return (size_t)kGramCtas * 1024 * sizeof(double); }
int extract_v_gram(cudaStream_t st, double* vw, size_t ldv, const double* a, size_t lda, size_t m, size_t w, const double* tau,
                   double* smat, size_t lds, void* ws) {
    if (m == 0 || w == 0) return NA_OK;
    if (w > 32) { set_error("extract_v_gram: w > 32"); return NA_EINVAL; }
    const unsigned grid = (unsigned)std::min<size_t>(std::min<size_t>(kGramCtas, (size_t)ctx().sm_count), ceil_div(m, (size_t)128));
    double* part = static_cast<double*>(ws);
    extract_v_gram_kernel<<<grid, 256, 0, st>>>(vw, (long long)ldv, a, (long long)lda, (long long)m, (int)w, tau, part);
    NAB_LAUNCH_CHECK();
    gram_finish_kernel<<<1, 1024, 0, st>>>(smat, (long long)lds, (int)w, tau, part, (int)grid);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// S = T^-1 = triu(V^T V, 1) + diag(1/tau)   (compact-WY identity T^-1 + T^-T = V^T V), in place on
// the Gram matrix g (w x w).  tau_j == 0 -> S_jj = 1 (the column of V is zero, so row/col j of g is 0).
// The strict lower triangle is zeroed so the block inverses of the TRSM see a clean triangle.
__global__ void build_s_kernel(double* __restrict__ g, long long ldg, int w, const double* __restrict__ tau, double tau_scale) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < w * w; idx += gridDim.x * blockDim.x) {
        const int i = idx % w, j = idx / w;
        if (i > j) g[i + j * ldg] = 0.0;
        else if (i == j) { const double t = tau[j] * tau_scale; g[i + j * ldg] = (t != 0.0) ? 1.0 / t : 1.0; }
    }
}
int build_s(cudaStream_t st, double* g, size_t ldg, size_t w, const double* tau) {
    if (w == 0) return NA_OK;
    build_s_kernel<<<(int)std::min<size_t>(ceil_div(w * w, 256), 256), 256, 0, st>>>(g, (long long)ldg, (int)w, tau, 1.0);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// tau_out[j] = (diag[j] != 0) ? 2 : 0   (nalgebra axes are unit vectors: H = I - 2 u u^T)
__global__ void tau_from_diag_kernel(double* __restrict__ tau_out, const double* __restrict__ diag, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) tau_out[i] = (diag[i] != 0.0) ? 2.0 : 0.0;
}
int tau_from_diag(cudaStream_t st, double* tau_out, const double* diag, size_t n) {
    if (n == 0) return NA_OK;
    tau_from_diag_kernel<<<(int)std::min<size_t>(ceil_div(n, 256), 64), 256, 0, st>>>(tau_out, diag, (int)n);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Storage conversion (classical -> nalgebra) and sign bookkeeping.
// csign[i] = c_i for i = 0..k (c_0 = 1; c_{i+1} = sign(beta_i) when reflected else c_i).
// ------------------------------------------------------------------------------------------------
// One CTA: beta_i / tau_i are fetched by all threads a chunk at a time (a single thread walking the diagonal pays one
// L2 round trip per column: 850 us at k = 4096), thread 0 runs the sign recurrence over the chunk in shared memory,
// then all threads form diag.
constexpr int kSignChunk = 2048;
// Columns [j0, j0 + k): the sign entering column j0 is csign[j0] (1 for j0 == 0), left there by the previous call.
__global__ void __launch_bounds__(1024) qr_signs_from_beta_kernel(const double* __restrict__ a, long long lda, const double* __restrict__ tau, int j0, int k,
                                                                  double* __restrict__ csign, double* __restrict__ diag) {
    __shared__ double beta_s[kSignChunk];
    __shared__ signed char refl_s[kSignChunk];
    __shared__ double c_s[kSignChunk + 1];
    __shared__ double carry;
    if (threadIdx.x == 0) {
        if (j0 == 0) csign[0] = 1.0;
        carry = j0 == 0 ? 1.0 : csign[j0];
    }
    __syncthreads();
    for (int i0 = j0; i0 < j0 + k; i0 += kSignChunk) {
        const int n = min(kSignChunk, j0 + k - i0);
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            beta_s[t] = a[(i0 + t) + (long long)(i0 + t) * lda];
            refl_s[t] = tau[i0 + t] != 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry;
            c_s[0] = c;
            for (int t = 0; t < n; ++t) {
                if (refl_s[t]) c = (beta_s[t] < 0.0) ? -1.0 : 1.0;
                c_s[t + 1] = c;
            }
            carry = c;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            diag[i0 + t] = refl_s[t] ? c_s[t] * beta_s[t] : 0.0;
            csign[i0 + t + 1] = c_s[t + 1];
        }
        __syncthreads();
    }
}
// csign from nalgebra's diag: c_{i+1} = c_i * signum(diag_i) when diag_i != 0 else c_i.
__global__ void qr_signs_from_diag_kernel(const double* __restrict__ diag, int k, double* __restrict__ csign) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double c = 1.0;
        csign[0] = c;
        for (int i = 0; i < k; ++i) {
            if (diag[i] != 0.0) c *= (diag[i] < 0.0) ? -1.0 : 1.0;
            csign[i + 1] = c;
        }
    }
}
// grid = (row blocks of 2048, columns): eight rows per thread, loads before stores, no 64-bit divisions
__global__ void __launch_bounds__(256) qr_convert_kernel(double* __restrict__ a, long long lda, long long m, long long j0, int k,
                                                         const double* __restrict__ tau, const double* __restrict__ csign, const double* __restrict__ diag) {
    const long long j = j0 + blockIdx.y;
    const long long i0 = (long long)blockIdx.x * 2048 + threadIdx.x;
    double* col = a + j * lda;
    const bool has_axis = j < k;
    const double t = has_axis ? tau[j] : 0.0;
    const double f = has_axis && t != 0.0 ? ((diag[j] < 0.0) ? 1.0 : -1.0) * sqrt(0.5 * t) : 0.0;      // -sign(diag_j)*sqrt(tau/2)
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const long long i = i0 + 256 * u; v[u] = i < m ? col[i] : 0.0; }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const long long i = i0 + 256 * u;
        if (i >= m) continue;
        if (i >= j) {                      // axis part of column j
            if (!has_axis) continue;
            col[i] = (t == 0.0) ? 0.0 : (i == j ? f : f * v[u]);
        } else {                           // strict upper: row i < k
            col[i] = csign[i + 1] * v[u];
        }
    }
}
// Columns [j0, j0 + ncols) of the classical storage -> nalgebra's, all rows.  The columns must be final, and so must
// csign[0 .. min(j0 + ncols, k)] -- call this panel by panel, left to right (each call extends csign / diag by the
// reflectors among its columns).  k = min(m, n) of the whole matrix.
int qr_convert_columns(cudaStream_t st, double* a, size_t lda, size_t m, size_t k, size_t j0, size_t ncols, const double* tau, double* csign,
                       double* diag) {
    if (ncols == 0 || m == 0) return NA_OK;
    if (j0 < k) {
        qr_signs_from_beta_kernel<<<1, 1024, 0, st>>>(a, (long long)lda, tau, (int)j0, (int)(std::min(k, j0 + ncols) - j0), csign, diag);
        NAB_LAUNCH_CHECK();
    }
    for (size_t c0 = j0; c0 < j0 + ncols; c0 += 65535) {            // gridDim.y limit
        const size_t ncol = std::min<size_t>(65535, j0 + ncols - c0);
        qr_convert_kernel<<<dim3((unsigned)ceil_div(m, (size_t)2048), (unsigned)ncol), 256, 0, st>>>(
            a, (long long)lda, (long long)m, (long long)c0, (int)k, tau, csign, diag);
        NAB_LAUNCH_CHECK();
    }
    return NA_OK;
}
int qr_convert_to_nalgebra(cudaStream_t st, double* a, size_t lda, size_t m, size_t n, const double* tau, double* csign, double* diag) {
    return qr_convert_columns(st, a, lda, m, std::min(m, n), 0, n, tau, csign, diag);
}
int qr_signs_from_diag(cudaStream_t st, const double* diag, size_t k, double* csign) {
    qr_signs_from_diag_kernel<<<1, 32, 0, st>>>(diag, (int)k, csign);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// b(i, j) *= csign[min(i, k-1) + 1]  (rows) or q(i, j) *= csign[j + 1] (cols)
__global__ void scale_signs_kernel(double* __restrict__ b, long long ldb, long long rows, long long cols, const double* __restrict__ csign,
                                   int k, int by_cols) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % rows, j = idx / rows;
        const double c = by_cols ? csign[j + 1] : csign[min(i, (long long)k - 1) + 1];
        if (c < 0.0) b[i + j * ldb] = -b[i + j * ldb];
    }
}
int scale_signs(cudaStream_t st, double* b, size_t ldb, size_t rows, size_t cols, const double* csign, size_t k, bool by_cols) {
    if (rows == 0 || cols == 0 || k == 0) return NA_OK;
    scale_signs_kernel<<<(int)std::min<size_t>(ceil_div(rows * cols, 256), (size_t)ctx().sm_count * 16), 256, 0, st>>>(
        b, (long long)ldb, (long long)rows, (long long)cols, csign, (int)k, by_cols ? 1 : 0);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab

extern "C" __attribute__((visibility("default"))) int na_debug_geqr2_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, nab::g_geqr2_prof, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(nab::g_geqr2_prof, z, sizeof(z)); }
    return 0;
}
