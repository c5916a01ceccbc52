// panel_lu.cu -- LU panel kernels: cooperative GETF2 with partial pivoting, swap-sequence ->
// permutation, parallel row permutation (LASWP).
//
// Reference semantics: LU::new's column step (/root/reference/src/linalg/lu.rs:103-119):
//   piv = icamax(A[i.., i]) + i   -- largest |x|, LOWEST index wins ties, a NaN only wins at
//                                    index 0 (src/base/min_max.rs:221-240)
//   diag == 0 -> skip the column (no swap, no scaling)               (lu.rs:107-110)
//   swap whole rows, multiply the column by 1/diag (reciprocal, lu.rs:344-349), rank-1 update.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "panel_lu.cuh"

namespace nab {

// ------------------------------------------------------------------------------------------------
// GETF2: m x w panel (w <= 128), rows distributed over G co-resident CTAs and kept in shared
// memory for the whole panel; one grid-wide barrier per column.
//
// Per column c every CTA publishes its local pivot candidate (|value|, row, the candidate's full
// row of w values) to a global slot; CTA 0 also publishes the current row c.  After the barrier
// every CTA reduces the G candidates to the same winner, reads the winner's row (= the pivot row)
// and row c, performs its share of the swap in shared memory, scales its rows of column c by the
// reciprocal pivot and applies the rank-1 update to columns c+1..w-1.  Slots are double-buffered
// by column parity (a CTA can be at most one barrier ahead).
// ------------------------------------------------------------------------------------------------
struct Getf2Params {
    double* a; long long lda;      // panel origin = A[j0, j0]
    int m, w;                      // panel rows / cols
    int rp;                        // rows per CTA (multiple of 32)
    int j0;                        // global row/col offset of the panel (for ipiv values)
    int* ipiv;                     // ipiv[j0 + c] = global pivot row of column c
    double2* xch;                  // [2][G][w + 2] (value, seq) pairs: |candidate|, candidate row, candidate's row values
    double2* rowc;                 // [2][w]        (value, seq) pairs: current row c, published by its owner
    int seq0;                      // sequence numbers already consumed in this workspace
};

__device__ long long g_getf2_prof[16];
#ifdef NAB_GETF2_PROF   // per-phase cycle counters of CTA 0 / thread 0 (tools/lu_timing.py), off in the product build
#define PROF(i) do { if (tid == 0 && cta == 0) { const long long t_ = clock64(); g_getf2_prof[i] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define PROF(i) do { (void)t_prev; } while (0)
#endif

// Software-pipelined column loop.  For column c every CTA
//   1. receives the candidates of column c (headers -> identical reduction in every CTA -> the
//      winner's published row = pivot row, and the old row c),
//   2. swaps, scales column c by the reciprocal pivot,
//   3. updates ONLY column c+1, finds its local pivot candidate and publishes it at once -- together
//      with that candidate's row (and row c+1), whose not-yet-updated entries are produced on the fly
//      with exactly the arithmetic step 4 will apply to them --
//   4. and only then applies the rank-1 update to columns c+2.. : the all-to-all exchange of column
//      c+1 travels through L2 while this bulk update runs.
__global__ void __launch_bounds__(256, 1) getf2_coop_kernel(const Getf2Params p) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int w = p.w, rp = p.rp, SL = w + 2;
    const int r_begin = cta * rp;
    const int nrows = max(0, min(rp, p.m - r_begin));
    double* s = sm;                        // [w][rp] column-major chunk
    double* prow = sm + (size_t)w * rp;    // [w] pivot row
    double* crow = prow + w;               // [w] old row c
    __shared__ double red_v[8];
    __shared__ int red_r[8], red_w[8];
    __shared__ int s_lrow;

    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) s[r + c * rp] = p.a[(long long)(r_begin + r) + (long long)c * p.lda];
    __syncthreads();
    const int ncol = min(w, p.m);
    long long t_prev = clock64();

    // Local winner of (bv, br) over the CTA -> s_lrow; publishes the header of column `col`.
    auto reduce_and_publish_header = [&](double bv, int br, int col) {
        int tag0 = 0;
        warp_best(bv, br, tag0);
        if (lane == 0) { red_v[warp] = bv; red_r[warp] = br; }
        __syncthreads();
        if (warp == 0) {
            double v = lane < 8 ? red_v[lane] : -2.0; int r = lane < 8 ? red_r[lane] : 0x7fffffff;
            warp_best(v, r, tag0);
            if (lane == 0) {
                if (G > 1) lu_st_pair(p.xch + ((size_t)(col & 1) * G + cta) * SL, v, pack_seq_row(p.seq0 + col + 1, r));
                s_lrow = r;
            }
        }
        __syncthreads();
    };

    // ---- prologue: candidates of column 0 (no pending update) ----
    {
        double bv = -1.0; int br = 0x7fffffff;
        for (int r = tid; r < nrows; r += nt) {
            const int gr = r_begin + r;
            const double v = pivot_key(s[r], gr == 0);
            if (cand_better(v, gr, bv, br)) { bv = v; br = gr; }
        }
        reduce_and_publish_header(bv, br, 0);
        if (G > 1) {
            const double seq = (double)(p.seq0 + 1);
            const int lr = s_lrow;
            double2* myslot = p.xch + (size_t)cta * SL;
            if (lr != 0x7fffffff)
                for (int cc = tid; cc < w; cc += nt) lu_st_pair(myslot + 2 + cc, s[(lr - r_begin) + cc * rp], seq);
            if (r_begin == 0 && nrows > 0)
                for (int cc = tid; cc < w; cc += nt) lu_st_pair(p.rowc + cc, s[cc * rp], seq);
        }
    }

    for (int c = 0; c < ncol; ++c) {
        const int par = c & 1;
        const int iseq = p.seq0 + c + 1;
        const double seq = (double)iseq;
        PROF(0);
        // ---- 1. receive column c ----
        int grow, gcta;
        if (G == 1) {
            grow = s_lrow; gcta = 0;
            if (grow != 0x7fffffff)
                for (int cc = tid; cc < w; cc += nt) { prow[cc] = s[grow + cc * rp]; crow[cc] = s[c + cc * rp]; }
            __syncthreads();
        } else {
            double gv = -2.0; grow = 0x7fffffff; gcta = -1;
            for (int i = tid; i < G; i += nt) {
                const double2* slot = p.xch + ((size_t)par * G + i) * SL;
                double sv, packed;
                do { lu_ld_pair_raw(slot, sv, packed); } while ((int)(__double_as_longlong(packed) >> 32) != iseq);
                const int sr = (int)(__double_as_longlong(packed) & 0xffffffffLL);
                if (cand_better(sv, sr, gv, grow)) { gv = sv; grow = sr; gcta = i; }
            }
            PROF(1);
            const int nw_used = min(8, (G + 31) / 32);
            if (warp < nw_used) {
                warp_best(gv, grow, gcta);
                if (lane == 0) { red_v[warp] = gv; red_r[warp] = grow; red_w[warp] = gcta; }
            }
            __syncthreads();
            gv = red_v[0]; grow = red_r[0]; gcta = red_w[0];
            for (int i = 1; i < nw_used; ++i)
                if (cand_better(red_v[i], red_r[i], gv, grow)) { gv = red_v[i]; grow = red_r[i]; gcta = red_w[i]; }
            PROF(2);
            const double2* wv = p.xch + ((size_t)par * G + gcta) * SL + 2;
            const double2* rc = p.rowc + par * w;
            for (int cc = tid; cc < w; cc += nt) {
                double a0, s0, a1, s1;
                lu_ld_pair_raw(wv + cc, a0, s0); lu_ld_pair_raw(rc + cc, a1, s1);
                while (s0 != seq) lu_ld_pair_raw(wv + cc, a0, s0);
                while (s1 != seq) lu_ld_pair_raw(rc + cc, a1, s1);
                prow[cc] = a0; crow[cc] = a1;
            }
            __syncthreads();
        }
        PROF(3);
        const double pivot = prow[c];
        const bool elim = pivot != 0.0;                  // lu.rs:107-110: an all-zero column is skipped
        if (cta == 0 && tid == 0) p.ipiv[p.j0 + c] = p.j0 + (elim ? grow : c);
        // ---- 2. swap rows c <-> grow inside shared memory, scale column c ----
        if (elim && grow != c) {
            const bool own_g = grow >= r_begin && grow < r_begin + nrows;
            const bool own_c = c >= r_begin && c < r_begin + nrows;
            if (own_g) for (int cc = tid; cc < w; cc += nt) s[(grow - r_begin) + cc * rp] = crow[cc];
            if (own_c) for (int cc = tid; cc < w; cc += nt) s[(c - r_begin) + cc * rp] = prow[cc];
            if (own_g || own_c) __syncthreads();
        }
        const int rlo = max(0, c + 1 - r_begin);         // first local row below the pivot row
        double* lc = s + (size_t)c * rp;
        if (elim) {
            const double inv = __drcp_rn(pivot);         // IEEE-rounded 1/diag, as gauss_step computes it
            for (int r = rlo + tid; r < nrows; r += nt) lc[r] = __dmul_rn(lc[r], inv);
            __syncthreads();
        }
        PROF(4);
        if (c + 1 >= w) continue;                        // last column of the panel: nothing right of it
        // ---- 3. column c+1 only: update, local candidate, publish header + candidate row + row c+1 ----
        {
            double* col = s + (size_t)(c + 1) * rp;
            const double pv = -prow[c + 1];
            double bv = -1.0; int br = 0x7fffffff;
            for (int r = rlo + tid; r < nrows; r += nt) {
                double v = col[r];
                if (elim) { v = __dadd_rn(__dmul_rn(pv, lc[r]), v); col[r] = v; }
                const int gr = r_begin + r;
                const double key = pivot_key(v, gr == c + 1);
                if (cand_better(key, gr, bv, br)) { bv = key; br = gr; }
            }
            if (c + 1 < ncol) {
                reduce_and_publish_header(bv, br, c + 1);
                if (G > 1) {
                    const double nseq = (double)(p.seq0 + c + 2);
                    const int npar = (c + 1) & 1;
                    const int lr = s_lrow;
                    double2* myslot = p.xch + ((size_t)npar * G + cta) * SL;
                    if (lr != 0x7fffffff) {
                        const int lrl = lr - r_begin;
                        const double ll = lc[lrl];
                        for (int cc = tid; cc < w; cc += nt) {
                            double v = s[lrl + cc * rp];
                            if (elim && cc >= c + 2) v = __dadd_rn(__dmul_rn(-prow[cc], ll), v);    // what step 4 will store
                            lu_st_pair(myslot + 2 + cc, v, nseq);
                        }
                    }
                    if (c + 1 >= r_begin && c + 1 < r_begin + nrows) {
                        const int rl = c + 1 - r_begin;
                        const double ll = lc[rl];
                        for (int cc = tid; cc < w; cc += nt) {
                            double v = s[rl + cc * rp];
                            if (elim && cc >= c + 2) v = __dadd_rn(__dmul_rn(-prow[cc], ll), v);
                            lu_st_pair(p.rowc + npar * w + cc, v, nseq);
                        }
                    }
                    __syncthreads();     // the rows above were read un-updated: step 4 may only start now
                }
            } else {
                __syncthreads();
            }
        }
        PROF(5);
        // ---- 4. bulk rank-1 update of columns c+2.. (unfused mul/add like the reference): warp `warp`
        //         owns the columns cc = c+2+warp (mod 8); a lane owns rows lane, lane+32, ... in chunks of
        //         8 whose multipliers stay in registers; the 8 row updates are independent ----
        if (elim) {
            for (int rb = (rlo / 256) * 256; rb < nrows; rb += 256) {
                double l[8];
                bool ok[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rb + lane + 32 * i;
                    ok[i] = r >= rlo && r < nrows;
                    l[i] = ok[i] ? lc[r] : 0.0;
                }
                for (int cc = c + 2 + warp; cc < w; cc += 8) {
                    double* col = s + (size_t)cc * rp;
                    const double pv = -prow[cc];
                    double v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = ok[i] ? col[rb + lane + 32 * i] : 0.0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __dadd_rn(__dmul_rn(pv, l[i]), v[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) if (ok[i]) col[rb + lane + 32 * i] = v[i];
                }
            }
        }
        __syncthreads();
        PROF(6);
    }
    __syncthreads();
    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) p.a[(long long)(r_begin + r) + (long long)c * p.lda] = s[r + c * rp];
}

// Factors the m x w panel at A[j0.., j0..j0+w).  ws: zero-initialised device workspace of
// getf2_workspace_bytes(); *seq_state (host) carries the sequence numbers consumed so far in it.
constexpr size_t kGetf2MaxCtas = 160;
size_t getf2_workspace_bytes() { return (2 * kGetf2MaxCtas * (kLuPanel + 2) + 2 * kLuPanel) * sizeof(double2); }

int getf2_panel(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, size_t j0, int* ipiv, void* ws, int* seq_state,
                int cta_limit) {
    if (m == 0 || w == 0) return NA_OK;
    if (w > (size_t)kLuPanel) { set_error("getf2: panel too wide"); return NA_EINVAL; }
    const int sms = ctx().sm_count;
    const size_t smem_budget = 200 * 1024;
    size_t rp_max = (smem_budget - 2 * w * 8) / (w * 8);
    rp_max = rp_max / 32 * 32;
    size_t G = ceil_div(m, rp_max);
    if (G > (size_t)sms) { set_error("getf2: panel of %zu x %zu rows does not fit %d SMs of shared memory", m, w, sms); return NA_EINVAL; }
    // spread rows evenly; use more CTAs (up to 64) when rows are plentiful
    size_t G_pref = std::min<size_t>(std::min<size_t>(64, (size_t)sms), ceil_div(m, (size_t)128));
    // cta_limit: SMs the caller keeps free for this cooperative launch while a persistent GEMM holds the rest;
    // a grid that does not fit would wait for that GEMM to finish (measured: 5.4 ms instead of 4.0 ms per panel)
    if (cta_limit > 0) G_pref = std::min<size_t>(G_pref, (size_t)cta_limit);
    if (G_pref > G) G = G_pref;
    size_t rp = round_up(ceil_div(m, G), 32);
    G = ceil_div(m, rp);
    const size_t smem = ((size_t)w * rp + 2 * w) * sizeof(double);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(getf2_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    Getf2Params p;
    p.a = a_panel; p.lda = (long long)lda; p.m = (int)m; p.w = (int)w; p.rp = (int)rp; p.j0 = (int)j0; p.ipiv = ipiv;
    p.xch = static_cast<double2*>(ws);
    p.rowc = p.xch + 2 * kGetf2MaxCtas * (kLuPanel + 2);
    p.seq0 = *seq_state;
    *seq_state += (int)w + 2 + ((w & 1) ? 1 : 0);      // keep the parity of seq0 even so buffers alternate cleanly
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)getf2_coop_kernel, dim3((unsigned)G), dim3(256), args, smem, st));
    count_launch();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// swap sequence -> permutation.  Applies swaps (a_s, b_s), s = 0..K-1, to the identity arrangement
// of rows [0, n) and emits the touched rows: dest[i] <- src[i] means "row dest[i] of the result is
// row src[i] of the input".  One CTA; the arrangement lives in shared memory (int32, n <= 51200)
// or in a global scratch array for taller matrices.
// ------------------------------------------------------------------------------------------------
// sa == nullptr: the swap sequence is (row0 + s, row0 + sb[s]) -- a panel-relative LAPACK-style ipiv (0-based) whose
// panel starts at row `row0`; otherwise the pairs (sa[s * stride], sb[s * stride]) and row0 = 0.
__global__ void __launch_bounds__(1024, 1)
perm_from_swaps_kernel(const int* __restrict__ sa, const int* __restrict__ sb, int K, int sa_stride, int row0, int n,
                       int* __restrict__ gperm, int use_global, int swaps_in_smem,
                       int* __restrict__ dest, int* __restrict__ src, int* __restrict__ count) {
    extern __shared__ int sperm[];
    int* perm = use_global ? gperm : sperm;
    // The swap list is staged in shared memory (when it fits) so that the serial simulation below never
    // waits on global memory; only the rows the list names are initialised and compacted (<= 2K of n).
    int* sx = sperm + (use_global ? 0 : n);
    int* sy = sx + K;
    const int tid = threadIdx.x, nt = blockDim.x;
    __shared__ int s_count;
    if (tid == 0) s_count = 0;
    for (int s = tid; s < K; s += nt) {
        const int x = sa ? sa[(size_t)s * sa_stride] : row0 + s, y = row0 + sb[(size_t)s * sa_stride];
        if (swaps_in_smem) { sx[s] = x; sy[s] = y; }
        perm[x] = x; perm[y] = y;                  // racing writers store the same value
    }
    __syncthreads();
    if (tid == 0) {
        if (swaps_in_smem) {
            for (int s = 0; s < K; ++s) {
                const int x = sx[s], y = sy[s];
                if (x != y) { const int t = perm[x]; perm[x] = perm[y]; perm[y] = t; }
            }
        } else {
            for (int s = 0; s < K; ++s) {
                const int x = sa ? sa[(size_t)s * sa_stride] : row0 + s, y = row0 + sb[(size_t)s * sa_stride];
                if (x != y) { const int t = perm[x]; perm[x] = perm[y]; perm[y] = t; }
            }
        }
    }
    __syncthreads();
    // compact the touched rows (order irrelevant); a row named several times is claimed once: the claim
    // resets its entry to the identity, so later claimants see an untouched row
    for (int s = tid; s < 2 * K; s += nt) {
        const int h = s >= K ? 1 : 0, si = s - h * K;
        const int r = swaps_in_smem ? (h ? sy[si] : sx[si])
                                    : (h ? row0 + sb[(size_t)si * sa_stride] : (sa ? sa[(size_t)si * sa_stride] : row0 + si));
        const int v = atomicExch(&perm[r], r);
        if (v != r) { const int i = atomicAdd(&s_count, 1); dest[i] = r; src[i] = v; }
    }
    __syncthreads();
    if (tid == 0) *count = s_count;
}

// rows of columns [0, ncols): out-of-place gather through shared memory, one CTA per column.
__global__ void __launch_bounds__(256)
permute_rows_kernel(double* __restrict__ a, long long lda, const int* __restrict__ dest, const int* __restrict__ src,
                    const int* __restrict__ count, int max_stage) {
    extern __shared__ double stage[];
    const int np = min(*count, max_stage);   // np <= max_stage is guaranteed by the host
    double* acol = a + (long long)blockIdx.x * lda;
    for (int i = threadIdx.x; i < np; i += blockDim.x) stage[i] = acol[src[i]];
    __syncthreads();
    for (int i = threadIdx.x; i < np; i += blockDim.x) acol[dest[i]] = stage[i];
}

// Variant for very long permutations: stage through global memory (tmp: np x ncols).
__global__ void permute_rows_gather_kernel(const double* __restrict__ a, long long lda, const int* __restrict__ src,
                                           const int* __restrict__ count, double* __restrict__ tmp, long long ldt, long long ncols) {
    const int np = *count;
    const long long total = (long long)np * ncols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % np, col = idx / np;
        tmp[i + col * ldt] = a[src[i] + col * lda];
    }
}
__global__ void permute_rows_scatter_kernel(double* __restrict__ a, long long lda, const int* __restrict__ dest,
                                            const int* __restrict__ count, const double* __restrict__ tmp, long long ldt, long long ncols) {
    const int np = *count;
    const long long total = (long long)np * ncols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % np, col = idx / np;
        a[dest[i] + col * lda] = tmp[i + col * ldt];
    }
}

// RowPerm workspace: dest[n], src[n], count, (gperm[n] when n is large)
size_t rowperm_workspace_bytes(size_t n) { return (3 * n + 64) * sizeof(int); }

// Builds the permutation equivalent to the swap sequence (sa[s*stride], sb[s*stride]), s < K, over rows [0, n).
static int rowperm_build_impl(cudaStream_t st, const int* sa, const int* sb, size_t K, size_t stride, int row0, size_t n, void* ws);
int rowperm_build(cudaStream_t st, const int* sa, const int* sb, size_t K, size_t stride, size_t n, void* ws) {
    return rowperm_build_impl(st, sa, sb, K, stride, 0, n, ws);
}
// The permutation of the swap sequence (row0 + s, row0 + ipiv[s]), s < K (ipiv: device, 0-based, relative to row0).
int rowperm_build_ipiv(cudaStream_t st, const int* ipiv, size_t K, int row0, size_t n, void* ws) {
    return rowperm_build_impl(st, nullptr, ipiv, K, 1, row0, n, ws);
}
static int rowperm_build_impl(cudaStream_t st, const int* sa, const int* sb, size_t K, size_t stride, int row0, size_t n, void* ws) {
    int* w = static_cast<int*>(ws);
    int* count = w; int* dest = w + 64; int* src = dest + n; int* gperm = src + n;
    const size_t budget = 200 * 1024;
    const bool use_global = n * sizeof(int) > budget;
    const size_t perm_bytes = use_global ? 0 : n * sizeof(int);
    const bool swaps_in_smem = perm_bytes + 2 * K * sizeof(int) <= budget;
    const size_t smem = perm_bytes + (swaps_in_smem ? 2 * K * sizeof(int) : 0);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(perm_from_swaps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    perm_from_swaps_kernel<<<1, 1024, smem, st>>>(sa, sb, (int)K, (int)stride, row0, (int)n, gperm, use_global ? 1 : 0, swaps_in_smem ? 1 : 0,
                                                   dest, src, count);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// Applies a built permutation to columns [0, ncols) of `a` (n rows).  max_touched bounds *count.
int rowperm_apply(cudaStream_t st, double* a, size_t lda, size_t ncols, size_t max_touched, const void* ws, size_t n) {
    const int* w = static_cast<const int*>(ws);
    return rowperm_apply_lists(st, a, lda, ncols, max_touched, w, w + 64, w + 64 + n);
}

// The same with explicit lists: row dest[i] of the result is row src[i] of the input, i < *count <= max_touched.
int rowperm_apply_lists(cudaStream_t st, double* a, size_t lda, size_t ncols, size_t max_touched, const int* count, const int* dest,
                        const int* src) {
    if (ncols == 0 || max_touched == 0) return NA_OK;
    const size_t stage_bytes = max_touched * sizeof(double);
    if (stage_bytes <= 200 * 1024) {
        static std::once_flag once;
        std::call_once(once, [] { cudaFuncSetAttribute(permute_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
        // one CTA per column; gridDim.x limited to 2^31-1
        permute_rows_kernel<<<(unsigned)ncols, 256, stage_bytes, st>>>(a, (long long)lda, dest, src, count, (int)max_touched);
        NAB_LAUNCH_CHECK();
        return NA_OK;
    }
    // global staging, in column chunks of <= 64 MiB
    const size_t chunk = std::max<size_t>(1, (64ull << 20) / stage_bytes);
    Scratch tmp;
    NAB_TRY(tmp.alloc(std::min(chunk, ncols) * stage_bytes, st));
    for (size_t c0 = 0; c0 < ncols; c0 += chunk) {
        const size_t nc = std::min(chunk, ncols - c0);
        const int blocks = ctx().sm_count * 8;
        permute_rows_gather_kernel<<<blocks, 256, 0, st>>>(a + c0 * lda, (long long)lda, src, count, tmp.as<double>(), (long long)max_touched, (long long)nc);
        NAB_LAUNCH_CHECK();
        permute_rows_scatter_kernel<<<blocks, 256, 0, st>>>(a + c0 * lda, (long long)lda, dest, count, tmp.as<double>(), (long long)max_touched, (long long)nc);
        NAB_LAUNCH_CHECK();
    }
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Direct small TRSM for the LU panel recursion: B <- L^-1 B with L the n1 x n1 UNIT lower triangle
// stored in `l` (n1 <= 128), B n1 x nrhs column-major, in place.  A CTA handles 64 columns of B with the
// current 64 x 64 block of L in shared memory; every substitution step is spread over all 256 threads
// (4 row groups x 64 columns), the entries of B stay in registers.  Replaces the TRTRI + GEMM + copy sequence (3 launches, 75-110 us) on the
// latency-critical panel chain by one launch of a few microseconds.
// ------------------------------------------------------------------------------------------------
constexpr int kTrsmCols = 64;      // columns of B per CTA
constexpr int kTrsmLdb = 65;       // row stride of the B tiles in shared memory (odd: conflict-free column access)

__device__ __forceinline__ void load_l_block(double* sl, const double* __restrict__ l, long long ldl, int r0, int c0, int n1,
                                             bool diag_block) {
    // sl(i, k) = L(r0 + i, c0 + k); outside the matrix or (diag_block) on/above the diagonal: 0
    for (int idx = threadIdx.x; idx < 64 * 64; idx += blockDim.x) {
        const int i = idx & 63, k = idx >> 6;
        const int r = r0 + i, c = c0 + k;
        double v = 0.0;
        if (r < n1 && c < n1 && (!diag_block || i > k)) v = l[r + (long long)c * ldl];
        sl[i + k * 64] = v;
    }
}
// Forward substitution with the unit-lower 64 x 64 block sl on a 64-column tile of B.  Thread (c = tid % 64,
// g = tid / 64) keeps rows g, g+4, ..., g+60 of column c in registers x[16]; at step k the owner of row k puts
// x_k into xs[c] (shared), one barrier, and every thread applies L(i, k) * x_k to its rows below k.  The k loop
// is fully unrolled so that all register indices are static.
// (No __restrict__ on the shared-memory pointers: xs is rewritten by other threads between barriers, and with
// restrict nvcc kept the value of an earlier step in a register across __syncthreads.)
__device__ __forceinline__ void trsm64_regs(const double* sl, volatile double* xs, double (&x)[16], int c, int g) {
#pragma unroll
    for (int k = 0; k < 63; ++k) {
        if (g == (k & 3)) xs[(k & 1) * 64 + c] = x[k >> 2];
        __syncthreads();
        const double xk = xs[(k & 1) * 64 + c];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            if (4 * t + 3 <= k) continue;                    // rows 4t .. 4t+3 are all <= k: nothing to do (static)
            const int i = g + 4 * t;
            if (i > k) x[t] = fma(-sl[i + k * 64], xk, x[t]);
        }
    }
}

__global__ void __launch_bounds__(256) trsm_unit_lower_small_kernel(const double* __restrict__ l, long long ldl, int n1,
                                                                   double* __restrict__ b, long long ldb, int nrhs) {
    extern __shared__ double tsm[];
    double* sl = tsm;                               // 64 x 64 block of L
    double* xs = sl + 64 * 64;                      // x_k of the current step, double-buffered by step parity
    double* xt = xs + 2 * 64;                       // solved top block (n1 > 64): xt[c * 65 + k]
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int col = blockIdx.x * kTrsmCols + c;
    const bool active = col < nrhs;
    double* bc = b + (long long)(active ? col : 0) * ldb;
    double x[16];
    load_l_block(sl, l, ldl, 0, 0, n1, true);
#pragma unroll
    for (int t = 0; t < 16; ++t) x[t] = (active && g + 4 * t < n1) ? bc[g + 4 * t] : 0.0;
    __syncthreads();
    trsm64_regs(sl, xs, x, c, g);
    if (active) {
#pragma unroll
        for (int t = 0; t < 16; ++t) if (g + 4 * t < n1) bc[g + 4 * t] = x[t];
    }
    if (n1 <= 64) return;
    // bottom rows: B_bot -= L21 * X_top, then the bottom-right triangle
#pragma unroll
    for (int t = 0; t < 16; ++t) xt[c * kTrsmLdb + g + 4 * t] = x[t];
    __syncthreads();
    load_l_block(sl, l, ldl, 64, 0, n1, false);
#pragma unroll
    for (int t = 0; t < 16; ++t) x[t] = (active && 64 + g + 4 * t < n1) ? bc[64 + g + 4 * t] : 0.0;
    __syncthreads();
    for (int k = 0; k < 64; ++k) {
        const double xk = xt[c * kTrsmLdb + k];
#pragma unroll
        for (int t = 0; t < 16; ++t) x[t] = fma(-sl[g + 4 * t + k * 64], xk, x[t]);
    }
    __syncthreads();
    load_l_block(sl, l, ldl, 64, 64, n1, true);
    __syncthreads();
    trsm64_regs(sl, xs, x, c, g);
    if (active) {
#pragma unroll
        for (int t = 0; t < 16; ++t) if (64 + g + 4 * t < n1) bc[64 + g + 4 * t] = x[t];
    }
}

int trsm_unit_lower_small(cudaStream_t st, size_t n1, const double* l, size_t ldl, double* b, size_t ldb, size_t nrhs) {
    if (n1 == 0 || nrhs == 0) return NA_OK;
    if (n1 > 128) { set_error("trsm_unit_lower_small: n1 > 128"); return NA_EINVAL; }
    const size_t smem = (64 * 64 + 2 * 64 + 64 * kTrsmLdb) * sizeof(double);
    static std::once_flag once;
    std::call_once(once, [smem] { cudaFuncSetAttribute(trsm_unit_lower_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
    trsm_unit_lower_small_kernel<<<(unsigned)ceil_div(nrhs, (size_t)kTrsmCols), 256, smem, st>>>(l, (long long)ldl, (int)n1, b, (long long)ldb, (int)nrhs);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ipiv (int, device) helpers -------------------------------------------------------------------
__global__ void iota_kernel(int* p, int n, int offset) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i + offset;
}
int iota_int(cudaStream_t st, int* p, size_t n, int offset) {
    if (n == 0) return NA_OK;
    iota_kernel<<<(int)std::min<size_t>(ceil_div(n, 256), 1024), 256, 0, st>>>(p, (int)n, offset);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab

#ifdef NAB_DEBUG_HOOKS   // debug build only (see nalgebra_b200/build.py): never part of the product library
extern "C" __attribute__((visibility("default"))) int na_debug_getf2_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, nab::g_getf2_prof, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(nab::g_getf2_prof, z, sizeof(z)); }
    return 0;
}

// test hook for tools/trsm_small_check.py (not part of include/nalgebra_b200.h)
extern "C" __attribute__((visibility("default"))) int na_debug_trsm_unit_lower_small(size_t n1, const double* l, size_t ldl, double* b,
                                                                                   size_t ldb, size_t nrhs, void* stream) {
    return nab::trsm_unit_lower_small(static_cast<cudaStream_t>(stream), n1, l, ldl, b, ldb, nrhs);
}
#endif  // NAB_DEBUG_HOOKS
