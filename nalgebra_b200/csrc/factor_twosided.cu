// factor_twosided.cu -- nalgebra's two-sided Householder reductions: Hessenberg, SymmetricTridiagonal, Bidiagonal.
//
// Reference: Hessenberg::new_with_workspace (/root/reference/src/linalg/hessenberg.rs:61-100) ->
// householder::clear_column_unchecked(.., shift = 1, bilateral) (src/linalg/householder.rs:61-85) ->
// reflection_axis_mut (householder.rs:19-53), Reflection::reflect_rows_with_sign / reflect_with_sign
// (src/geometry/reflection.rs:70-131); SymmetricTridiagonal::new (src/linalg/symmetric_tridiagonal.rs:54-95: hegemv,
// dotc, three hegerc, src/base/blas.rs:359-420, 868-900); Bidiagonal::new (src/linalg/bidiagonal.rs:74-150:
// clear_column_unchecked / clear_row_unchecked, householder.rs:92-127).
//
// The reference applies every reflector with Level-1 sweeps: a product pass and an update pass over the trailing
// matrix per side (four to six trips through memory per step).  These reductions cannot be blocked without changing
// which vectors get stored, so they stay one-reflector-per-step, memory-bound Level-2 work -- but one persistent
// cooperative kernel per factorization does each step in the minimum number of passes over HBM:
//   Hessenberg / SymmetricTridiagonal: ONE read pass that forms both products at once (w = A u along rows and
//     z = A^T u along columns of the same tile), ONE read-modify-write pass that applies both sides at once
//     (d_j = u . (s A + (-2 s u_j) w)_j is s z_j + (-2 s u_j)(u . w) by linearity, so no pass over the half-updated
//     matrix is needed);
//   Bidiagonal: read pass (column products), read pass (row products of the column-reflected matrix formed on the
//     fly), one read-modify-write pass that applies both reflections.
// The trailing matrix is cut into (row block) x (column chunk) tasks, one per CTA: a thread owns U rows (coalesced
// along the column), walks the columns of its chunk, keeps its row products in registers and reduces column products
// with warp shuffles.  Partial products go to small per-task arrays and are summed in a fixed order (run-to-run
// deterministic, no atomics).  The last CTA is the "leader": while the others run the update pass it updates the next
// pivot column alone and turns it into the next reflection axis, so axis construction is off the critical path; two
// grid barriers per step (four for Bidiagonal).  Per-element update arithmetic is the reference's (unfused multiply,
// then add, same operand order); only the order of summation inside the products differs (results to rounding).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace nab {

namespace ts {
// -DNAB_TS_PROF: nanoseconds (globaltimer) spent by CTA 0 / the leader in each phase, summed over the steps
#ifdef NAB_TS_PROF
__device__ unsigned long long g_ts_prof[16];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TS_MARK(slot)                                                                  \
    do {                                                                               \
        if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {    \
            const unsigned long long now__ = gtime();                                  \
            g_ts_prof[(blockIdx.x == 0 ? 0 : 8) + (slot)] += now__ - ts_t0;            \
            ts_t0 = now__;                                                             \
        }                                                                              \
    } while (0)
#define TS_START() unsigned long long ts_t0 = gtime()
#else
#define TS_MARK(slot) do {} while (0)
#define TS_START() do {} while (0)
#endif
constexpr int T = 512;            // threads per CTA (16 warps: the dependent FP64 chains of an element need the other warps to hide)
constexpr int U = 2;              // rows per thread
constexpr int RB = T * U;         // rows per row block
constexpr int NCB_MAX = 64;       // column chunks per row block (= partial row products to sum)
constexpr int CB = 256;           // columns whose per-column factors are staged in shared memory at a time
constexpr int NW = T / 32;
constexpr int CG = 8;              // columns a thread has in flight (x U rows)
constexpr int ZB = 64;             // columns between two CTA-wide reductions of the column products

enum Mode { HESS = 0, SYM = 1, BD_Z = 2, BD_W = 3 };

struct Params {
    double* a; long long lda; int m, n;
    double* d;        // Hessenberg: subdiag; SymmetricTridiagonal: off_diagonal; Bidiagonal: diagonal
    double* e;        // Bidiagonal: off_diagonal
    double* wpart;    // [NCB_MAX][m]   partial row products by column chunk
    double* wfull;    // [m]            their sums, written by the last task of a row block to finish the read pass
    unsigned* cnt;    // [ceil(m / RB)] tasks of a row block that have delivered their partials (back to 0 after each pass)
    double* zpart;    // [ceil(m / RB)][n] partial column products by row block
    double* gpart;    // [G] partial u . w by task
    double* fvec;     // [n] Bidiagonal: column-reflection factors of the step
    double* vvec;     // [n] Bidiagonal: the row axis, contiguous
    double* hh;       // [2][4] (sign, reflected) of the column axis and of the row axis, by step parity
    long long set2;   // fused kernel: offset (in doubles) of the second copy of wpart / wfull / zpart / gpart (step parity)
};
// the partial-product arrays of step parity q
__device__ __forceinline__ Params with_parity(const Params& p, int q) {
    Params r = p;
    const long long o = q ? p.set2 : 0;
    r.wpart += o; r.wfull += o; r.zpart += o; r.gpart += o;
    return r;
}

struct Smem {
    double zs[2][ZB][NW];
    double c1[CB], c2[CB];
    double red[NW];
    double bc[4];
    int last;
};

// ---- tiling of rows [row0, row0 + nrows) x columns [col0, col0 + ncols); tri: only j - col0 <= r - row0 ------------
struct Tiling {
    int row0, nrows, col0, ncols, nrb, gt, total; bool tri;      // 32-bit: gt * width and the summed widths stay below 2^31 for any matrix that fits the HBM
    __device__ int width(int rb) const { return tri ? min(ncols, min(nrows, (rb + 1) * RB)) : ncols; }
    __device__ int ncb(int rb) const {
        const int w = width(rb);
        int c = (int)((unsigned)(gt * w) / (unsigned)total);
        c = min(c, NCB_MAX);
        c = min(c, (w + 63) / 64);
        return max(c, 1);
    }
    __device__ void init(int gt_, int row0_, int nrows_, int col0_, int ncols_, bool tri_) {
        gt = gt_; row0 = row0_; nrows = nrows_; col0 = col0_; ncols = ncols_; tri = tri_;
        nrb = (nrows + RB - 1) / RB;
        total = 0;
        for (int rb = 0; rb < nrb; ++rb) total += width(rb);
        if (total < 1) total = 1;
    }
    __device__ int ntasks() const { int t = 0; for (int rb = 0; rb < nrb; ++rb) t += ncb(rb); return t; }
    __device__ int rb_of_row(int r) const { return (r - row0) / RB; }
};
struct Task { int rb, cb, ncb, r0, r1, j0, j1; };
__device__ inline bool find_task(const Tiling& tl, int t, Task& tk) {
    int acc = 0;
    for (int rb = 0; rb < tl.nrb; ++rb) {
        const int c = tl.ncb(rb);
        if (t < acc + c) {
            tk.rb = rb; tk.cb = t - acc; tk.ncb = c;
            tk.r0 = tl.row0 + rb * RB; tk.r1 = min(tl.row0 + tl.nrows, tk.r0 + RB);
            const int w = tl.width(rb);
            int cw = (w + c - 1) / c; cw = (cw + CG - 1) & ~(CG - 1);
            tk.j0 = tl.col0 + min(w, tk.cb * cw); tk.j1 = tl.col0 + min(w, (tk.cb + 1) * cw);
            return true;
        }
        acc += c;
    }
    return false;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += red[w];
    return t;
}

// householder::reflection_axis_mut on a contiguous vector, by one CTA.  Returns the value the reference returns
// (-signed_norm when a reflection is needed, signed_norm otherwise); *reflected says which.
template <int R = 16>                                            // up to R entries per thread stay in registers: one trip through memory
__device__ double make_axis(double* x, int len, double* red, bool* reflected) {
    const int tid = threadIdx.x;
    const bool in_regs = len <= R * T;
    double xr[R];
    double s = 0.0;
    if (in_regs) {
#pragma unroll
        for (int u = 0; u < R; ++u) { const int r = tid + u * T; xr[u] = r < len ? x[r] : 0.0; }
#pragma unroll
        for (int u = 0; u < R; ++u) s = fma(xr[u], xr[u], s);
    } else {
        for (int r = tid; r < len; r += T) { const double v = x[r]; s = fma(v, v, s); }
    }
    const double sq = block_sum(s, red);
    const double nrm = sqrt(sq);
    const double x0 = x[0];
    const double modulus = x0 >= 0.0 ? x0 : -x0, sign = x0 >= 0.0 ? 1.0 : -1.0;       // simba to_exp
    const double signed_norm = sign * nrm;
    const double factor = (sq + modulus * nrm) * 2.0;
    __syncthreads();                                             // x[0] read by everyone before it changes
    if (factor != 0.0) {
        const double sf = sqrt(factor);
        double s2 = 0.0;
        if (in_regs) {
            if (tid == 0) xr[0] = x0 + signed_norm;
#pragma unroll
            for (int u = 0; u < R; ++u) { xr[u] = xr[u] / sf; s2 = fma(xr[u], xr[u], s2); }      // unscale_mut
        } else {
            for (int r = tid; r < len; r += T) { const double v = (r == 0 ? x0 + signed_norm : x[r]) / sf; s2 = fma(v, v, s2); }
        }
        const double nn = sqrt(block_sum(s2, red));              // the second normalisation (householder.rs:38-46)
        if (in_regs) {
#pragma unroll
            for (int u = 0; u < R; ++u) { const int r = tid + u * T; if (r < len) x[r] = xr[u] / nn; }
        } else {
            for (int r = tid; r < len; r += T) { const double v = (r == 0 ? x0 + signed_norm : x[r]) / sf; x[r] = v / nn; }
        }
        *reflected = true;
        return -signed_norm;
    }
    if (tid == 0) x[0] = x0 + signed_norm;
    *reflected = false;
    return signed_norm;
}
__device__ __forceinline__ double signum_of(double v) { return signbit(v) ? -1.0 : 1.0; }

// Partial sums are fetched eight at a time (independent L2 loads in flight) and added in index order.
__device__ __forceinline__ double sum_strided(const double* base, size_t stride, int count) {
    double acc = 0.0;
    for (int q0 = 0; q0 < count; q0 += 8) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = q0 + q < count ? __ldcg(base + (size_t)(q0 + q) * stride) : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += v[q];
    }
    return acc;
}
// sum of the column-product partials of column j: row blocks b0 .. nrb - 1
__device__ __forceinline__ double sum_zpart(const Params& p, const Tiling& tl, int j) {
    const int b0 = tl.tri ? (j - tl.col0) / RB : 0;
    return sum_strided(p.zpart + (size_t)b0 * p.n + j, (size_t)p.n, tl.nrb - b0);
}
__device__ __forceinline__ double sum_gpart(const Params& p, int ntasks, Smem& sm) {
    double g = 0.0;
    if (threadIdx.x < 32) {
        for (int t = threadIdx.x; t < ntasks; t += 32) g += __ldcg(p.gpart + t);
        g = warp_sum(g);
        if (threadIdx.x == 0) sm.bc[0] = g;
    }
    __syncthreads();
    g = sm.bc[0];
    __syncthreads();
    return g;
}

// ---- the read pass of one task: row products w_r = sum_j e(r, j) x_j (registers -> wpart) and column products
// z_j = sum_r e(r, j) y_r (-> zpart), plus the task's share of y . w (-> gpart).
//   HESS: e = a, x_j = u_j, y_r = u_r (0 above the axis);   SYM: lower triangle, w over j <= r, z over r > j;
//   BD_Z: z only, y_r = u_r;   BD_W: w only, e = column-reflected element f_j u_r + s a formed on the fly, x_j = v_j.
template <int MODE>
__device__ void pass_reduce(const Params& p, const Tiling& tl, const Task& tk, int task_id, int k, double su, bool refl_u, Smem& sm) {
    constexpr bool WANT_W = MODE != BD_Z, WANT_Z = MODE != BD_W, TRI = MODE == SYM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lda = p.lda;
    const double* ucol = p.a + (long long)k * lda;               // the column axis lives in column k
    int r[U]; bool ok[U]; double y[U], wacc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        r[i] = tk.r0 + tid + T * i; ok[i] = r[i] < tk.r1;
        y[i] = (ok[i] && (MODE != HESS || r[i] > k)) ? ucol[r[i]] : 0.0;
        wacc[i] = 0.0;
    }
    double gacc = 0.0;
    int it = 0;
    for (int jb = tk.j0; jb < tk.j1; jb += ZB, ++it) {
        const int buf = it & 1;
        for (int jj = 0; jj < ZB && jb + jj < tk.j1; jj += CG) {
            // CG columns x U rows per thread in flight
            double av[CG][U], x[CG], f[CG];
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                const int j = jb + jj + q;
                const bool jok = j < tk.j1;
                x[q] = 0.0; f[q] = 0.0;
                if (WANT_W && jok) x[q] = MODE == BD_W ? __ldcg(p.vvec + j) : ucol[j];
                if (MODE == BD_W && jok) f[q] = __ldcg(p.fvec + j);
#pragma unroll
                for (int i = 0; i < U; ++i) av[q][i] = (ok[i] && jok && (!TRI || j <= r[i])) ? p.a[r[i] + (long long)j * lda] : 0.0;
            }
            double zl[CG];
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                const int j = jb + jj + q;
                zl[q] = 0.0;
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    double ev = av[q][i];
                    if (MODE == BD_W && refl_u) ev = ok[i] ? __dadd_rn(__dmul_rn(f[q], y[i]), __dmul_rn(su, ev)) : 0.0;
                    if (WANT_W) wacc[i] = fma(ev, x[q], wacc[i]);
                    if (WANT_Z) zl[q] = (!TRI || r[i] > j) ? fma(ev, y[i], zl[q]) : zl[q];
                }
            }
            if (WANT_Z) {
                // eight column sums across the warp in 4 + 2 + 1 + 1 + 1 exchanges: each round a lane keeps half of its
                // values and hands the other half to its partner (fixed tree: deterministic)
                static_assert(CG == 8, "the exchange pattern below is written for eight columns");
                double h4[4], h2[2], h1;
                const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double send = b16 ? zl[q] : zl[q + 4], keep = b16 ? zl[q + 4] : zl[q];
                    h4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double send = b8 ? h4[q] : h4[q + 2], keep = b8 ? h4[q + 2] : h4[q];
                    h2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                {
                    const double send = b4 ? h2[0] : h2[1], keep = b4 ? h2[1] : h2[0];
                    h1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
                h1 += __shfl_xor_sync(0xffffffffu, h1, 2);
                h1 += __shfl_xor_sync(0xffffffffu, h1, 1);
                if ((lane & 3) == 0) sm.zs[buf][jj + (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0)][warp] = h1;
            }
        }
        if (WANT_Z) {
            __syncthreads();
            if (tid < ZB && jb + tid < tk.j1) {
                double z = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) z += sm.zs[buf][tid][w];
                p.zpart[(size_t)tk.rb * p.n + jb + tid] = z;
                if (MODE == HESS || MODE == SYM) gacc = fma(ucol[jb + tid], z, gacc);
            }
        }
    }
    if (WANT_W) {
#pragma unroll
        for (int i = 0; i < U; ++i)
            if (ok[i]) { p.wpart[(size_t)tk.cb * p.m + r[i]] = wacc[i]; if (MODE == SYM) gacc = fma(y[i], wacc[i], gacc); }
        // the last task of this row block to get here sums the partials (in chunk order, whoever does it) into wfull
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned done = atomicAdd(p.cnt + tk.rb, 1u);
            sm.last = done + 1 == (unsigned)tk.ncb;
            if (sm.last) p.cnt[tk.rb] = 0;
            __threadfence();
        }
        __syncthreads();
        if (sm.last) {
#pragma unroll
            for (int i = 0; i < U; ++i)
                if (ok[i]) p.wfull[r[i]] = sum_strided(p.wpart + r[i], (size_t)p.m, tk.ncb);
        }
    }
    if (MODE == HESS || MODE == SYM) {
        const double g = block_sum(gacc, sm.red);
        if (tid == 0) p.gpart[task_id] = g;
    }
    __syncthreads();                                             // zs / red are free for the next task
}

// ---- the read-modify-write pass of one task -------------------------------------------------------------------------
//   HESS: a1 = (m2s u_j) w_r + s a;  rows below the pivot row: a2 = F_j u_r + s a1, F_j = (s z_j + (m2s u_j) g) m2s
//   SYM : a = (-u_j) p_r + a;  a = (-p_j) u_r + a;  a = ((dot 2) u_j) u_r + a          (rows >= columns only)
//   BD  : a1 = f_j u_r + su a (if the column reflected);  a2 = (m2sv v_j) w_r + sv a1 (if the row reflected)
// skip_col: the column the leader updates itself.
template <int MODE>
__device__ void pass_update(const Params& p, const Tiling& tl, const Task& tk, int k, int skip_col, double su, bool refl_u, double sv,
                            bool refl_v, double g, Smem& sm) {
    const int tid = threadIdx.x;
    const long long lda = p.lda;
    const double* ucol = p.a + (long long)k * lda;
    const double m2su = su * -2.0, m2sv = sv * -2.0;
    int r[U]; bool ok[U]; double ur[U], wr[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        r[i] = tk.r0 + tid + T * i; ok[i] = r[i] < tk.r1;
        ur[i] = (ok[i] && (MODE != HESS || r[i] > k)) ? ucol[r[i]] : 0.0;
        wr[i] = 0.0;
        if (ok[i]) {
            if (MODE == SYM) wr[i] = 2.0 * (__ldcg(p.wfull + r[i]) + sum_zpart(p, tl, r[i]));        // p_r
            else if (MODE == HESS || refl_v) wr[i] = __ldcg(p.wfull + r[i]);
        }
    }
    for (int jb = tk.j0; jb < tk.j1; jb += CB) {
        const int nb = min(CB, tk.j1 - jb);
        __syncthreads();
        for (int c = tid; c < nb; c += T) {
            const int j = jb + c;
            if (MODE == HESS) {
                const double c1 = m2su * ucol[j];
                sm.c1[c] = c1;
                sm.c2[c] = __dmul_rn(__dadd_rn(__dmul_rn(su, sum_zpart(p, tl, j)), __dmul_rn(c1, g)), m2su);
            } else if (MODE == SYM) {
                sm.c1[c] = ucol[j];
                sm.c2[c] = 2.0 * (__ldcg(p.wfull + j) + sum_zpart(p, tl, j));                             // p_j
            } else {
                sm.c1[c] = refl_u ? __ldcg(p.fvec + j) : 0.0;
                sm.c2[c] = refl_v ? m2sv * __ldcg(p.vvec + j) : 0.0;
            }
        }
        __syncthreads();
        for (int c0 = 0; c0 < nb; c0 += CG) {
            double av[CG][U];
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                const int j = jb + c0 + q;
#pragma unroll
                for (int i = 0; i < U; ++i)
                    av[q][i] = (ok[i] && c0 + q < nb && j != skip_col && (MODE != SYM || j <= r[i])) ? p.a[r[i] + (long long)j * lda] : 0.0;
            }
            // branch-free per element (masked entries were loaded as zeros and are kept out by the predicated store)
#pragma unroll
            for (int q = 0; q < CG; ++q) {
                const int j = jb + c0 + q;
                const bool cin = c0 + q < nb && j != skip_col;
                const double c1 = sm.c1[c0 + q], c2 = sm.c2[c0 + q];
#pragma unroll
                for (int i = 0; i < U; ++i) {
                    const bool act = ok[i] && cin && (MODE != SYM || j <= r[i]);
                    double v = av[q][i];
                    if (MODE == HESS) {
                        const double v1 = __dadd_rn(__dmul_rn(c1, wr[i]), __dmul_rn(su, v));
                        const double v2 = __dadd_rn(__dmul_rn(c2, ur[i]), __dmul_rn(su, v1));
                        v = r[i] > k ? v2 : v1;
                    } else if (MODE == SYM) {
                        v = __dadd_rn(__dmul_rn(-c1, wr[i]), v);
                        v = __dadd_rn(__dmul_rn(-c2, ur[i]), v);
                        v = __dadd_rn(__dmul_rn(__dmul_rn(g, c1), ur[i]), v);           // g = dot * 2
                    } else {
                        if (refl_u) v = __dadd_rn(__dmul_rn(c1, ur[i]), __dmul_rn(su, v));
                        if (refl_v) v = __dadd_rn(__dmul_rn(c2, wr[i]), __dmul_rn(sv, v));
                    }
                    if (act) p.a[r[i] + (long long)j * lda] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Hessenberg (SYMM = false) and SymmetricTridiagonal (SYMM = true): axes in column k, rows k + 1..
// ------------------------------------------------------------------------------------------------------------------
template <bool SYMM>
__global__ void __launch_bounds__(T, 1) two_sided_kernel(const Params p) {
    constexpr int MODE = SYMM ? SYM : HESS;
    cg::grid_group grid = cg::this_grid();
    __shared__ Smem sm;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x, Gt = G - 1;
    const bool leader = cta == G - 1;
    const int n = p.n;
    const long long lda = p.lda;
    if (leader) {
        bool refl;
        const double nr = make_axis(p.a + 1, n - 1, sm.red, &refl);
        if (tid == 0) { p.d[0] = nr; p.hh[0] = signum_of(nr); p.hh[1] = refl ? 1.0 : 0.0; }
    }
    grid.sync();
    TS_START();
    for (int k = 0; k + 1 < n; ++k) {
        const double* hh = p.hh + 4 * (k & 1);
        double* hn = p.hh + 4 * ((k + 1) & 1);
        const double su = __ldcg(hh + 0);
        const bool refl = __ldcg(hh + 1) != 0.0;
        const int c = k + 1;                                        // the leader's column
        Tiling tl;
        tl.init(Gt, SYMM ? k + 1 : 0, SYMM ? n - k - 1 : n, k + 1, n - k - 1, SYMM);
        const int ntasks = tl.ntasks();
        TS_MARK(0);
        if (refl) {
            if (!leader)
                for (int t = cta; t < ntasks; t += Gt) { Task tk; if (find_task(tl, t, tk)) pass_reduce<MODE>(p, tl, tk, t, k, su, true, sm); }
            TS_MARK(1);
            grid.sync();
            TS_MARK(2);
            double g = sum_gpart(p, ntasks, sm);
            if (SYMM) g = (2.0 * g) * 2.0;                           // dot = u . p = 2 u . (w + z); the reference uses dot * 2
            if (!leader) {
                for (int t = cta; t < ntasks; t += Gt) { Task tk; if (find_task(tl, t, tk)) pass_update<MODE>(p, tl, tk, k, c, su, true, 1.0, false, g, sm); }
            } else {
                // column c of the updated matrix, then the axis of step k + 1 out of its rows c + 1..
                const double* ucol = p.a + (long long)k * lda;
                double* col = p.a + (long long)c * lda;
                const double m2s = su * -2.0, uc = ucol[c];
                if (!SYMM) {
                    const double c1 = m2s * uc;
                    const double c2 = __dmul_rn(__dadd_rn(__dmul_rn(su, sum_zpart(p, tl, c)), __dmul_rn(c1, g)), m2s);
                    for (int r = tid; r < n; r += T) {
                        double v = __dadd_rn(__dmul_rn(c1, __ldcg(p.wfull + r)), __dmul_rn(su, col[r]));
                        if (r > k) v = __dadd_rn(__dmul_rn(c2, ucol[r]), __dmul_rn(su, v));
                        col[r] = v;
                    }
                } else {
                    const double pc = 2.0 * (__ldcg(p.wfull + c) + sum_zpart(p, tl, c));
                    for (int r = c + tid; r < n; r += T) {
                        const double pr = 2.0 * (__ldcg(p.wfull + r) + sum_zpart(p, tl, r));
                        double v = col[r];
                        v = __dadd_rn(__dmul_rn(-uc, pr), v);
                        v = __dadd_rn(__dmul_rn(-pc, ucol[r]), v);
                        v = __dadd_rn(__dmul_rn(__dmul_rn(g, uc), ucol[r]), v);
                        col[r] = v;
                    }
                }
                __syncthreads();
            }
            TS_MARK(3);
        }
        if (leader && c + 1 < n) {
            bool r2;
            const double nr = make_axis(p.a + (long long)c * lda + c + 1, n - c - 1, sm.red, &r2);
            if (tid == 0) { p.d[c] = nr; hn[0] = signum_of(nr); hn[1] = r2 ? 1.0 : 0.0; }
        }
        TS_MARK(4);
        grid.sync();
        TS_MARK(5);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused variant (the default): ONE pass per step.  The update of step s - 1 is applied to an element while it sits in a
// register and the products with the axis of step s are accumulated from the updated value before it is stored: one
// read + one write of the trailing block per step (16 n (n - k) / 8 (n - k)^2 bytes) instead of two reads + one write.
// The axis of step s must exist before the pass, so the leader's work (update of the next pivot column with the products
// of the pass that just ended, axis construction) sits between the two grid barriers of a step while the other CTAs wait
// (~8 us), which is cheaper than a second trip through memory from n ~ 2000 on.  Partial products are double-buffered by
// step parity: the pass of step s reads the sums of step s - 1 while it writes those of step s.
//   pend: an update (step s - 1; axis up = column s - 1, sign sup, scalar gp) is outstanding; prod: step s reflects.
// ------------------------------------------------------------------------------------------------------------------
template <int MODE>
__device__ void pass_fused(const Params& pr, const Params& pw, const Tiling& tlp, const Tiling& tl, const Task& tk, int task_id, int s,
                           bool pend, double sup, double gp, bool prod, Smem& sm, double* sx) {
    constexpr bool TRI = MODE == SYM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long lda = pr.lda;
    double* a = pr.a;
    const double* up = a + (long long)(s - 1) * lda;             // axis of the outstanding update (valid when pend)
    const double* un = a + (long long)s * lda;                   // axis of this step
    const double m2sp = sup * -2.0;
    int r[U]; bool ok[U]; double wr[U], ur[U], y[U], wacc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        r[i] = tk.r0 + tid + T * i; ok[i] = r[i] < tk.r1;
        wr[i] = 0.0; ur[i] = 0.0; wacc[i] = 0.0;
        if (ok[i] && pend) {
            if (MODE == SYM) wr[i] = 2.0 * (__ldcg(pr.wfull + r[i]) + sum_zpart(pr, tlp, r[i]));      // p_r of step s - 1
            else wr[i] = __ldcg(pr.wfull + r[i]);
            if (MODE != HESS || r[i] > s - 1) ur[i] = up[r[i]];
        }
        y[i] = (ok[i] && prod && (MODE != HESS || r[i] > s)) ? un[r[i]] : 0.0;
    }
    double gacc = 0.0;
    int it = 0;
    // Software pipeline over half-groups of four columns: the loads of the next half-group are issued before the current
    // one is updated / accumulated / stored, so a thread always has 16 - 32 loads in flight (the registers are the same
    // 8 x U doubles as one eight-column group).
    constexpr int HG = CG / 2;
    double bufA[HG][U], bufB[HG][U];
    auto load_half = [&](double (&bf)[HG][U], int j0) {
#pragma unroll
        for (int q = 0; q < HG; ++q) {
            const int j = j0 + q;
#pragma unroll
            for (int i = 0; i < U; ++i) bf[q][i] = (ok[i] && j < tk.j1 && (!TRI || j <= r[i])) ? a[r[i] + (long long)j * lda] : 0.0;
        }
    };
    // Branch-free: masked elements (rows beyond the block, columns beyond the chunk, the strict upper triangle of the
    // symmetric case) were loaded as zeros, go through the same arithmetic and are kept out by one predicated store and two
    // selects -- per-element `continue`s cost a divergence barrier pair each.
    bool left[U];
#pragma unroll
    for (int i = 0; i < U; ++i) left[i] = r[i] > s - 1;
    auto compute_half = [&](double (&bf)[HG][U], int jb, int c0, int nb, double* zl) {
#pragma unroll
        for (int q = 0; q < HG; ++q) {
            const int c = c0 + q, j = jb + c;
            const bool cin = c < nb;
            const double c1 = sm.c1[c], c2 = sm.c2[c], x = sx[c];      // c < CB always: stale entries beyond nb are never used
            double zq = 0.0;
#pragma unroll
            for (int i = 0; i < U; ++i) {
                const bool act = ok[i] && cin && (!TRI || j <= r[i]);
                double v = bf[q][i];
                if (pend) {
                    if (MODE == HESS) {
                        const double v1 = __dadd_rn(__dmul_rn(c1, wr[i]), __dmul_rn(sup, v));
                        const double v2 = __dadd_rn(__dmul_rn(c2, ur[i]), __dmul_rn(sup, v1));
                        v = left[i] ? v2 : v1;
                    } else {
                        v = __dadd_rn(__dmul_rn(-c1, wr[i]), v);
                        v = __dadd_rn(__dmul_rn(-c2, ur[i]), v);
                        v = __dadd_rn(__dmul_rn(__dmul_rn(gp, c1), ur[i]), v);
                    }
                    if (act) a[r[i] + (long long)j * lda] = v;
                }
                const double ve = act ? v : 0.0;
                wacc[i] = fma(ve, x, wacc[i]);
                const double yz = (!TRI || r[i] > j) ? y[i] : 0.0;
                zq = fma(ve, yz, zq);
            }
            zl[q] = zq;
        }
    };
    load_half(bufA, tk.j0);
    for (int jb = tk.j0; jb < tk.j1; jb += CB) {
        const int nb = min(CB, tk.j1 - jb);
        __syncthreads();
        for (int c = tid; c < nb; c += T) {
            const int j = jb + c;
            double c1 = 0.0, c2 = 0.0;
            if (pend) {
                if (MODE == HESS) {
                    c1 = m2sp * up[j];
                    c2 = __dmul_rn(__dadd_rn(__dmul_rn(sup, sum_zpart(pr, tlp, j)), __dmul_rn(c1, gp)), m2sp);
                } else {
                    c1 = up[j];
                    c2 = 2.0 * (__ldcg(pr.wfull + j) + sum_zpart(pr, tlp, j));                           // p_j
                }
            }
            sm.c1[c] = c1; sm.c2[c] = c2;
            sx[c] = prod ? un[j] : 0.0;
        }
        __syncthreads();
        for (int z0 = 0; z0 < nb; z0 += ZB, ++it) {
            const int buf = it & 1;
            for (int jj = 0; jj < ZB && z0 + jj < nb; jj += CG) {
                const int c0 = z0 + jj;
                double zl[CG];
                load_half(bufB, jb + c0 + HG);
                compute_half(bufA, jb, c0, nb, zl);
                load_half(bufA, jb + c0 + CG);                       // the next group (possibly of the next batch: only addresses matter)
                compute_half(bufB, jb, c0 + HG, nb, zl + HG);
                if (prod) {
                    static_assert(CG == 8, "the exchange pattern below is written for eight columns");
                    double h4[4], h2[2], h1;
                    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double send = b16 ? zl[q] : zl[q + 4], keep = b16 ? zl[q + 4] : zl[q];
                        h4[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const double send = b8 ? h4[q] : h4[q + 2], keep = b8 ? h4[q + 2] : h4[q];
                        h2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    {
                        const double send = b4 ? h2[0] : h2[1], keep = b4 ? h2[1] : h2[0];
                        h1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    h1 += __shfl_xor_sync(0xffffffffu, h1, 2);
                    h1 += __shfl_xor_sync(0xffffffffu, h1, 1);
                    if ((lane & 3) == 0) sm.zs[buf][jj + (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0)][warp] = h1;
                }
            }
            if (prod) {
                __syncthreads();
                if (tid < ZB && z0 + tid < nb) {
                    double z = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) z += sm.zs[buf][tid][w];
                    pw.zpart[(size_t)tk.rb * pw.n + jb + z0 + tid] = z;
                    gacc = fma(sx[z0 + tid], z, gacc);
                }
            }
        }
    }
    if (prod) {
#pragma unroll
        for (int i = 0; i < U; ++i)
            if (ok[i]) { pw.wpart[(size_t)tk.cb * pw.m + r[i]] = wacc[i]; if (MODE == SYM) gacc = fma(y[i], wacc[i], gacc); }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned done = atomicAdd(pw.cnt + tk.rb, 1u);
            sm.last = done + 1 == (unsigned)tk.ncb;
            if (sm.last) pw.cnt[tk.rb] = 0;
            __threadfence();
        }
        __syncthreads();
        if (sm.last) {
#pragma unroll
            for (int i = 0; i < U; ++i)
                if (ok[i]) pw.wfull[r[i]] = sum_strided(pw.wpart + r[i], (size_t)pw.m, tk.ncb);
        }
        const double g = block_sum(gacc, sm.red);
        if (tid == 0) pw.gpart[task_id] = g;
    }
    __syncthreads();
}

template <bool SYMM>
__global__ void __launch_bounds__(T, 1) two_sided_fused_kernel(const Params p) {
    constexpr int MODE = SYMM ? SYM : HESS;
    cg::grid_group grid = cg::this_grid();
    __shared__ Smem sm;
    __shared__ double sx[CB];
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x, Gt = G - 1;
    const bool leader = cta == G - 1;
    const int n = p.n;
    const long long lda = p.lda;
    if (leader) {
        bool refl;
        const double nr = make_axis<16>(p.a + 1, n - 1, sm.red, &refl);
        if (tid == 0) { p.d[0] = nr; p.hh[0] = signum_of(nr); p.hh[1] = refl ? 1.0 : 0.0; }
    }
    grid.sync();
    TS_START();
    bool pend = false;
    double sup = 1.0, gp = 0.0;
    Tiling tlp;
    tlp.init(Gt, SYMM ? 1 : 0, SYMM ? n - 1 : n, 1, n - 1, SYMM);
    for (int s = 0; s + 1 < n; ++s) {
        const double* hh = p.hh + 4 * (s & 1);
        double* hn = p.hh + 4 * ((s + 1) & 1);
        const double su = __ldcg(hh + 0);
        const bool prod = __ldcg(hh + 1) != 0.0;
        const int c = s + 1;                                        // the leader's column
        Tiling tl;
        tl.init(Gt, SYMM ? s + 1 : 0, SYMM ? n - s - 1 : n, s + 1, n - s - 1, SYMM);
        const int ntasks = tl.ntasks();
        const Params pr = with_parity(p, (s + 1) & 1), pw = with_parity(p, s & 1);      // sums of step s - 1 / of step s
        TS_MARK(0);
        if ((pend || prod) && !leader)
            for (int t = cta; t < ntasks; t += Gt) { Task tk; if (find_task(tl, t, tk)) pass_fused<MODE>(pr, pw, tlp, tl, tk, t, s, pend, sup, gp, prod, sm, sx); }
        TS_MARK(1);
        grid.sync();
        TS_MARK(2);
        double g = 0.0;
        if (prod) {
            g = sum_gpart(pw, ntasks, sm);
            if (SYMM) g = (2.0 * g) * 2.0;                           // dot = u . p = 2 u . (w + z); the reference uses dot * 2
        }
        if (leader) {
            // column c carries every update up to step s - 1 (the pass just applied the last one): apply step s, then
            // turn its rows c + 1.. into the axis of step s + 1
            if (prod) {
                const double* ucol = p.a + (long long)s * lda;
                double* col = p.a + (long long)c * lda;
                const double m2s = su * -2.0, uc = ucol[c];
                // eight rows per thread at a time, every load of the batch issued before the first use (the leader sits
                // between the two barriers of the step: its latency is everybody's)
                constexpr int LB = 8;
                if (!SYMM) {
                    const double c1 = m2s * uc;
                    const double c2 = __dmul_rn(__dadd_rn(__dmul_rn(su, sum_zpart(pw, tl, c)), __dmul_rn(c1, g)), m2s);
                    for (int r0 = tid; r0 < n; r0 += LB * T) {
                        double wv[LB], cv[LB], uv[LB];
#pragma unroll
                        for (int q = 0; q < LB; ++q) {
                            const int r = r0 + q * T;
                            wv[q] = r < n ? __ldcg(pw.wfull + r) : 0.0; cv[q] = r < n ? col[r] : 0.0; uv[q] = (r < n && r > s) ? ucol[r] : 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < LB; ++q) {
                            const int r = r0 + q * T;
                            if (r >= n) continue;
                            double v = __dadd_rn(__dmul_rn(c1, wv[q]), __dmul_rn(su, cv[q]));
                            if (r > s) v = __dadd_rn(__dmul_rn(c2, uv[q]), __dmul_rn(su, v));
                            col[r] = v;
                        }
                    }
                } else {
                    const double pc = 2.0 * (__ldcg(pw.wfull + c) + sum_zpart(pw, tl, c));
                    for (int r0 = c + tid; r0 < n; r0 += LB * T) {
                        double wv[LB], cv[LB], uv[LB], zv[LB];
#pragma unroll
                        for (int q = 0; q < LB; ++q) {
                            const int r = r0 + q * T;
                            wv[q] = r < n ? __ldcg(pw.wfull + r) : 0.0; cv[q] = r < n ? col[r] : 0.0; uv[q] = r < n ? ucol[r] : 0.0; zv[q] = 0.0;
                        }
                        for (int b = 0; b < tl.nrb; ++b) {              // column products of row r: row blocks (r - col0) / RB ..
#pragma unroll
                            for (int q = 0; q < LB; ++q) {
                                const int r = r0 + q * T;
                                if (r < n && b >= (r - tl.col0) / RB) zv[q] += __ldcg(pw.zpart + (size_t)b * pw.n + r);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < LB; ++q) {
                            const int r = r0 + q * T;
                            if (r >= n) continue;
                            const double prr = 2.0 * (wv[q] + zv[q]);
                            double v = cv[q];
                            v = __dadd_rn(__dmul_rn(-uc, prr), v);
                            v = __dadd_rn(__dmul_rn(-pc, uv[q]), v);
                            v = __dadd_rn(__dmul_rn(__dmul_rn(g, uc), uv[q]), v);
                            col[r] = v;
                        }
                    }
                }
                __syncthreads();
            }
            TS_MARK(3);
            if (c + 1 < n) {
                bool r2;
                const double nr = make_axis<16>(p.a + (long long)c * lda + c + 1, n - c - 1, sm.red, &r2);
                if (tid == 0) { p.d[c] = nr; hn[0] = signum_of(nr); hn[1] = r2 ? 1.0 : 0.0; }
            }
        }
        TS_MARK(4);
        grid.sync();
        TS_MARK(5);
        pend = prod; sup = su; gp = g; tlp = tl;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Bidiagonal, m >= n (upper bidiagonal; the wide case runs on the transpose): column axes in column k rows k..,
// row axes in row k columns k + 1..
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T, 1) bidiagonal_kernel(const Params p) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Smem sm;
    const int tid = threadIdx.x, cta = blockIdx.x, G = gridDim.x, Gt = G - 1;
    const bool leader = cta == G - 1;
    const int m = p.m, n = p.n;
    const long long lda = p.lda;
    if (leader) {
        bool refl;
        const double nr = make_axis<8>(p.a, m, sm.red, &refl);
        if (tid == 0) { p.d[0] = nr; p.hh[0] = signum_of(nr); p.hh[1] = refl ? 1.0 : 0.0; }
    }
    grid.sync();
    for (int k = 0; k + 1 < n; ++k) {
        double* hh = p.hh + 4 * (k & 1);
        double* hn = p.hh + 4 * ((k + 1) & 1);
        const double su = __ldcg(hh + 0);
        const bool refl_u = __ldcg(hh + 1) != 0.0;
        const double* ucol = p.a + (long long)k * lda;
        const int c = k + 1;
        Tiling tz, tw;
        tz.init(Gt, k, m - k, k + 1, n - k - 1, false);             // column products: rows k.., columns k + 1..
        tw.init(Gt, k + 1, m - k - 1, k + 1, n - k - 1, false);     // row products / update: rows k + 1..
        // ---- pass 1: z_j = u . a_j
        if (refl_u && !leader) {
            const int nt = tz.ntasks();
            for (int t = cta; t < nt; t += Gt) { Task tk; if (find_task(tz, t, tk)) pass_reduce<BD_Z>(p, tz, tk, t, k, su, true, sm); }
        }
        grid.sync();
        // ---- leader: factors f_j, row k of the column-reflected matrix, the row axis v (stored in row k and in vvec)
        if (leader) {
            const double m2s = su * -2.0, u0 = ucol[k];
            for (int j = c + tid; j < n; j += T) {
                double v = p.a[k + (long long)j * lda];
                if (refl_u) {
                    const double f = sum_zpart(p, tz, j) * m2s;       // reflection.rs:79
                    p.fvec[j] = f;
                    v = __dadd_rn(__dmul_rn(f, u0), __dmul_rn(su, v));
                }
                p.vvec[j] = v;
            }
            __syncthreads();
            bool rv;
            const double nr = make_axis<8>(p.vvec + c, n - c, sm.red, &rv);
            __syncthreads();
            for (int j = c + tid; j < n; j += T) p.a[k + (long long)j * lda] = p.vvec[j];
            if (tid == 0) { p.e[k] = nr; hh[2] = signum_of(nr); hh[3] = rv ? 1.0 : 0.0; }
        }
        grid.sync();
        const double sv = __ldcg(hh + 2);
        const bool refl_v = __ldcg(hh + 3) != 0.0;
        // ---- pass 2: w_r = (column-reflected row r) . v
        if (refl_v && !leader) {
            const int nt = tw.ntasks();
            for (int t = cta; t < nt; t += Gt) { Task tk; if (find_task(tw, t, tk)) pass_reduce<BD_W>(p, tw, tk, t, k, su, refl_u, sm); }
        }
        grid.sync();
        // ---- pass 3: both reflections applied; the leader does column c and the next column axis
        if (refl_u || refl_v) {
            if (!leader) {
                const int nt = tw.ntasks();
                for (int t = cta; t < nt; t += Gt) { Task tk; if (find_task(tw, t, tk)) pass_update<BD_W>(p, tw, tk, k, c, su, refl_u, sv, refl_v, 0.0, sm); }
            } else {
                double* col = p.a + (long long)c * lda;
                const double fc = refl_u ? __ldcg(p.fvec + c) : 0.0;
                const double c2 = refl_v ? (sv * -2.0) * __ldcg(p.vvec + c) : 0.0;
                for (int r = c + tid; r < m; r += T) {
                    double v = col[r];
                    if (refl_u) v = __dadd_rn(__dmul_rn(fc, ucol[r]), __dmul_rn(su, v));
                    if (refl_v) v = __dadd_rn(__dmul_rn(c2, __ldcg(p.wfull + r)), __dmul_rn(sv, v));
                    col[r] = v;
                }
                __syncthreads();
            }
        }
        if (leader) {
            bool r2;
            const double nr = make_axis<8>(p.a + (long long)c * lda + c, m - c, sm.red, &r2);
            if (tid == 0) { p.d[c] = nr; hn[0] = signum_of(nr); hn[1] = r2 ? 1.0 : 0.0; }
        }
        grid.sync();
    }
}

static int two_sided_grid(size_t m, size_t n) {
    // one task per CTA plus the leader; small problems take fewer CTAs (cheaper grid barriers)
    const size_t tasks = ceil_div(m, (size_t)RB) * std::max<size_t>(1, std::min<size_t>(NCB_MAX, n / 64));
    return (int)std::min<size_t>((size_t)ctx().sm_count, std::max<size_t>(2, tasks + 1));
}

static int launch(const void* kernel, cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* d, double* e) {
    const int G = two_sided_grid(m, n);
    const size_t nrb = ceil_div(m, (size_t)RB);
    Scratch ws;
    const size_t ng = (size_t)G + nrb;                            // tasks: at most one per CTA plus one per row block
    const size_t set = (size_t)NCB_MAX * m + m + nrb * n + ng;     // wpart | wfull | zpart | gpart, two copies (step parity)
    const size_t words = 2 * set + 2 * n + 8 + nrb;
    NAB_TRY(ws.alloc(words * sizeof(double), s));
    NAB_CUDA(cudaMemsetAsync(ws.p, 0, words * sizeof(double), s));
    double* w = ws.as<double>();
    double* wf = w + (size_t)NCB_MAX * m; double* zp = wf + m; double* gp = zp + nrb * n; double* fv = w + 2 * set;
    Params p{a, (long long)lda, (int)m, (int)n, d, e, w, wf, reinterpret_cast<unsigned*>(fv + 2 * n + 8), zp, gp, fv, fv + n, fv + 2 * n,
             (long long)set};
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel(kernel, dim3((unsigned)G), dim3(T), args, 0, s));
    count_launch();
    return NA_OK;
}
// One fused pass per step (two trips through memory per element instead of three, but the leader's ~15 us between the two
// barriers of a step are exposed) pays once the trailing block is well beyond the L2: measured (profiles/r02_twosided_timing.txt)
// Hessenberg 6144 568 / 575 ms, 8192 1154 / 1271, 12288 3463 / 4019, 16384 7839 / 9337 (fused / two-pass);
// SymmetricTridiagonal 8192 699 / 644, 12288 1971 / 1971, 16384 4241 / 4334.  na_set_tuning("ts_fused", 0 | 1) forces
// one of them (tests), -1 restores the size rule; NAB_TS_FUSED=0|1 does the same from the environment.
static long g_ts_fused = [] { const char* e = getenv("NAB_TS_FUSED"); return e ? (long)(atoi(e) != 0) : -1L; }();
static bool ts_fused(size_t n, bool symmetric) {
    if (g_ts_fused >= 0) return g_ts_fused != 0;
    return n >= (symmetric ? (size_t)16384 : (size_t)6144);
}
}  // namespace ts
void ts_set_fused(long v) { ts::g_ts_fused = v < 0 ? -1 : (v != 0); }

// subdiag: DEVICE, n - 1 signed norms (nalgebra's Hessenberg::subdiag)
int hessenberg_device(cudaStream_t s, size_t n, double* a, size_t lda, double* subdiag) {
    if (n < 2) return NA_OK;
    if (lda < n) { set_error("hessenberg: lda < n"); return NA_EINVAL; }
    if (n > 0x7fffffull) { set_error("hessenberg: dimension exceeds 2^23"); return NA_EINVAL; }
    return ts::launch(ts::ts_fused(n, false) ? (const void*)ts::two_sided_fused_kernel<false> : (const void*)ts::two_sided_kernel<false>, s, n, n, a, lda, subdiag, nullptr);
}
// off_diagonal: DEVICE, n - 1 signed norms; only the lower triangle of a is read / written
int symmetric_tridiagonal_device(cudaStream_t s, size_t n, double* a, size_t lda, double* off_diagonal) {
    if (n < 2) return NA_OK;
    if (lda < n) { set_error("symmetric_tridiagonal: lda < n"); return NA_EINVAL; }
    if (n > 0x7fffffull) { set_error("symmetric_tridiagonal: dimension exceeds 2^23"); return NA_EINVAL; }
    return ts::launch(ts::ts_fused(n, true) ? (const void*)ts::two_sided_fused_kernel<true> : (const void*)ts::two_sided_kernel<true>, s, n, n, a, lda, off_diagonal, nullptr);
}
// diagonal (min(m, n)) / off_diagonal (min(m, n) - 1): DEVICE.  m < n runs on the transpose (the row step of the wide
// case is the column step of the tall one, householder.rs:92-127 against :61-85).
int bidiagonal_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal) {
    if (std::min(m, n) == 0) return NA_OK;
    if (lda < m) { set_error("bidiagonal: lda < m"); return NA_EINVAL; }
    if (m > 0x7fffffull || n > 0x7fffffull) { set_error("bidiagonal: dimension exceeds 2^23"); return NA_EINVAL; }
    if (m >= n) return ts::launch((const void*)ts::bidiagonal_kernel, s, m, n, a, lda, diagonal, off_diagonal);
    Scratch t;
    NAB_TRY(t.alloc(m * n * sizeof(double), s));
    NAB_TRY(copy_strided(s, t.as<double>(), 1, (ptrdiff_t)n, a, (ptrdiff_t)lda, 1, n, m));           // t (n x m) = a^T
    NAB_TRY(ts::launch((const void*)ts::bidiagonal_kernel, s, n, m, t.as<double>(), n, diagonal, off_diagonal));
    NAB_TRY(copy_strided(s, a, 1, (ptrdiff_t)lda, t.as<double>(), (ptrdiff_t)n, 1, m, n));
    return NA_OK;
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_hessenberg_f64_dev(size_t n, double* a, size_t lda, double* subdiag, void* stream) {
    NAB_TRY(ensure_init());
    if (n == 0) { set_error("Cannot compute the hessenberg decomposition of an empty matrix."); return NA_EINVAL; }
    if (!a || (n > 1 && !subdiag)) { set_error("hessenberg: null argument"); return NA_EINVAL; }
    return hessenberg_device(static_cast<cudaStream_t>(stream), n, a, lda, subdiag);
}

int na_symmetric_tridiagonal_f64_dev(size_t n, double* a, size_t lda, double* off_diagonal, void* stream) {
    NAB_TRY(ensure_init());
    if (n == 0) { set_error("Unable to compute the symmetric tridiagonal decomposition of an empty matrix."); return NA_EINVAL; }
    if (!a || (n > 1 && !off_diagonal)) { set_error("symmetric_tridiagonal: null argument"); return NA_EINVAL; }
    return symmetric_tridiagonal_device(static_cast<cudaStream_t>(stream), n, a, lda, off_diagonal);
}

int na_bidiagonal_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal, void* stream) {
    NAB_TRY(ensure_init());
    const size_t mn = std::min(m, n);
    if (mn == 0) { set_error("Cannot compute the bidiagonalization of an empty matrix."); return NA_EINVAL; }
    if (!a || !diagonal || (mn > 1 && !off_diagonal)) { set_error("bidiagonal: null argument"); return NA_EINVAL; }
    return bidiagonal_device(static_cast<cudaStream_t>(stream), m, n, a, lda, diagonal, off_diagonal);
}

// host-pointer twins: upload, reduce, download (the first step touches the whole matrix and every step rewrites the
// trailing block, so nothing can stream)
static int two_sided_host(int which, size_t m, size_t n, double* a, size_t lda, double* d, double* e) {
    const size_t mn = std::min(m, n);
    if (!a || lda < m) { set_error("two-sided reduction: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch da, dd; size_t ldd;
    NAB_TRY(upload_matrix(s, da, ldd, a, lda, m, n));
    NAB_TRY(dd.alloc(2 * std::max<size_t>(mn, 1) * sizeof(double), s));
    double* d_dev = dd.as<double>(); double* e_dev = d_dev + std::max<size_t>(mn, 1);
    size_t nd = 0, ne = 0;
    if (which == 0) { NAB_TRY(na_hessenberg_f64_dev(n, da.as<double>(), ldd, d_dev, s)); nd = n - 1; }
    else if (which == 1) { NAB_TRY(na_symmetric_tridiagonal_f64_dev(n, da.as<double>(), ldd, d_dev, s)); nd = n - 1; }
    else { NAB_TRY(na_bidiagonal_f64_dev(m, n, da.as<double>(), ldd, d_dev, e_dev, s)); nd = mn; ne = mn - 1; }
    NAB_TRY(download_matrix(s, a, lda, da.as<double>(), ldd, m, n));
    if (nd) NAB_CUDA(cudaMemcpyAsync(d, d_dev, nd * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (ne) NAB_CUDA(cudaMemcpyAsync(e, e_dev, ne * sizeof(double), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

int na_hessenberg_f64(size_t n, double* a, size_t lda, double* subdiag) {
    NAB_TRY(ensure_init());
    if (n == 0) { set_error("Cannot compute the hessenberg decomposition of an empty matrix."); return NA_EINVAL; }
    if (n > 1 && !subdiag) { set_error("hessenberg: null argument"); return NA_EINVAL; }
    return two_sided_host(0, n, n, a, lda, subdiag, nullptr);
}

int na_symmetric_tridiagonal_f64(size_t n, double* a, size_t lda, double* off_diagonal) {
    NAB_TRY(ensure_init());
    if (n == 0) { set_error("Unable to compute the symmetric tridiagonal decomposition of an empty matrix."); return NA_EINVAL; }
    if (n > 1 && !off_diagonal) { set_error("symmetric_tridiagonal: null argument"); return NA_EINVAL; }
    return two_sided_host(1, n, n, a, lda, off_diagonal, nullptr);
}

int na_bidiagonal_f64(size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal) {
    NAB_TRY(ensure_init());
    const size_t mn = std::min(m, n);
    if (mn == 0) { set_error("Cannot compute the bidiagonalization of an empty matrix."); return NA_EINVAL; }
    if (!diagonal || (mn > 1 && !off_diagonal)) { set_error("bidiagonal: null argument"); return NA_EINVAL; }
    return two_sided_host(2, m, n, a, lda, diagonal, off_diagonal);
}

#ifdef NAB_TS_PROF
// debug build only: out[16] = nanoseconds per phase (CTA 0: slots 0..7, leader: 8..15) since the last call; resets them
NAB_API int na_debug_ts_prof(unsigned long long* out) {
    NAB_CUDA(cudaDeviceSynchronize());
    NAB_CUDA(cudaMemcpyFromSymbol(out, ts::g_ts_prof, 16 * sizeof(unsigned long long)));
    unsigned long long zero[16] = {0};
    NAB_CUDA(cudaMemcpyToSymbol(ts::g_ts_prof, zero, sizeof(zero)));
    return NA_OK;
}
#endif

}  // extern "C"
