// factor_pivoted.cu -- the two factorizations of nalgebra that pivot on the largest entry of the whole trailing matrix:
// FullPivLU (P A Q = L U) and ColPivQR (A P = Q R with the pivot column chosen by icamax_full).
//
// Reference: FullPivLU::new (/root/reference/src/linalg/full_piv_lu.rs:56-91) -> icamax_full
// (src/base/min_max.rs:146-167: column-major scan, strict >, first maximum wins) -> swap_columns / swap_rows ->
// lu::gauss_step(_swap) (src/linalg/lu.rs:337-389); ColPivQR::new (src/linalg/col_piv_qr.rs:56-93) -> icamax_full ->
// swap_columns -> householder::clear_column_unchecked (src/linalg/householder.rs:19-85, geometry/reflection.rs:70-83).
//
// Both need the position of the largest |a_ij| of the trailing matrix before every step, so they cannot be blocked
// like LU / QR: every step is one pass over the trailing matrix (rank-1 update fused with the search for the next
// pivot), i.e. memory-bound Level-2 work -- L2-resident up to n ~ 3500, HBM beyond.  One persistent cooperative
// kernel per factorization: warps own columns (lanes along rows: coalesced), grid-wide barriers separate the phases of
// a step, every CTA reduces the per-CTA pivot candidates in the same order.  FullPivLU applies exactly the
// reference's arithmetic (IEEE reciprocal, unfused multiply then add), so its packed factors and both permutation
// sequences are bit-identical to the reference's; ColPivQR's dot products are summed in a different order (results to
// rounding, pivots identical unless two candidates are within rounding of each other).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace nab {

namespace pv {
#ifndef NAB_PV_T
#define NAB_PV_T 512
#endif
constexpr int RU = 12;             // rows per lane in flight in the column sweeps
constexpr int T = NAB_PV_T;      // warps own columns: the loads a CTA keeps in flight scale with its warp count

struct Cand { double key; long long pos; double val; };   // key: |value| (-1: can never win); pos = row + col * m (column-major order); val: the entry itself

// icamax_full's rule: strict > in a column-major scan = the largest key, the lowest position among equals.  NaN never
// compares greater, so it only "wins" when it is the first element scanned (then nothing replaces it): first_nan_key.
__device__ __forceinline__ double cand_key(double v, bool first) {
    const double a = fabs(v);
    if (a != a) return first ? __longlong_as_double(0x7ff0000000000000ll) : -1.0;      // NaN at the first position: +inf
    return a;
}
__device__ __forceinline__ bool cand_better(double k1, long long p1, double k2, long long p2) {
    return k1 > k2 || (k1 == k2 && p1 < p2);
}
__device__ __forceinline__ void warp_best(double& k, long long& p, double& v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double k2 = __shfl_xor_sync(0xffffffffu, k, o);
        const long long p2 = __shfl_xor_sync(0xffffffffu, p, o);
        const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
        if (cand_better(k2, p2, k, p)) { k = k2; p = p2; v = v2; }
    }
}
// CTA-wide winner -> cand[cta]
__device__ __forceinline__ void block_publish(double k, long long p, double v, Cand* cand, double* sk, long long* sp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sv = sk + T / 32;
    warp_best(k, p, v);
    if (lane == 0) { sk[warp] = k; sp[warp] = p; sv[warp] = v; }
    __syncthreads();
    if (warp == 0) {
        k = lane < T / 32 ? sk[lane] : -2.0; p = lane < T / 32 ? sp[lane] : 0x7fffffffffffffffll; v = lane < T / 32 ? sv[lane] : 0.0;
        warp_best(k, p, v);
        if (lane == 0) { cand[blockIdx.x].key = k; cand[blockIdx.x].pos = p; cand[blockIdx.x].val = v; }
    }
    __syncthreads();
}
// every CTA: the same winner out of the G candidates (after a grid barrier)
__device__ __forceinline__ long long grid_winner(const Cand* cand, int G, long long* s_pos, double* s_val) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        double k = -2.0, v = 0.0; long long p = 0x7fffffffffffffffll;
        for (int g = lane; g < G; g += 32) {
            const double k2 = __ldcg(&cand[g].key); const long long p2 = __ldcg(&cand[g].pos);
            if (cand_better(k2, p2, k, p)) { k = k2; p = p2; v = __ldcg(&cand[g].val); }
        }
        warp_best(k, p, v);
        if (lane == 0) { *s_pos = p; *s_val = v; }
    }
    __syncthreads();
    return *s_pos;
}
}  // namespace pv

struct PivotedParams {
    double* a; long long lda; int m, n;
    int* p_row;            // [min(m,n)] pivot row of step i (FullPivLU), i where nothing moved
    int* p_col;            // [min(m,n)] pivot column of step i
    double* diag;          // [min(m,n)] ColPivQR: signed norms
    pv::Cand* cand;        // [G]
    double* hh;            // ColPivQR: [2][2] (sign, reflected) of the current reflector by step parity, published by CTA 0
    int* steps_done;       // FullPivLU: number of elimination steps before an exactly zero pivot stopped it
};

// ------------------------------------------------------------------------------------------------------------------
// FullPivLU.  Two grid barriers per step:
//   phase S: the column swap i <-> pc and the row swap i <-> pr in one go (rows outside {i, pr} of the two columns,
//            columns outside {i, pc} of the two rows, and the four corner elements by one thread), plus the scaling of
//            the PREVIOUS pivot column by its reciprocal pivot (its unscaled entries were what the update read);
//   phase U: rank-1 update of the trailing matrix in the reference's arithmetic fused with the search for the next
//            pivot; a candidate carries its signed value, so the pivot never has to be read back from the matrix.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(pv::T, 1) full_piv_lu_kernel(const PivotedParams p) {
    using namespace pv;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sk[2 * (T / 32)];
    __shared__ long long sp[T / 32];
    __shared__ long long s_pos;
    __shared__ double s_val;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int m = p.m, n = p.n, mn = min(m, n);
    const long long lda = p.lda;
    double* a = p.a;
    const int gwarp = cta * (T / 32) + warp, nwarps = G * (T / 32);
    const long long gtid = (long long)cta * T + tid, nthreads = (long long)G * T;

    // pivot of step 0: the whole matrix
    {
        double bk = -2.0, bv = 0.0; long long bp = 0x7fffffffffffffffll;
        for (int j = gwarp; j < n; j += nwarps)
            for (int r = lane; r < m; r += 32) {
                const double x = a[r + j * lda];
                const double k = cand_key(x, r == 0 && j == 0);
                const long long pos = r + (long long)j * m;
                if (cand_better(k, pos, bk, bp)) { bk = k; bp = pos; bv = x; }
            }
        block_publish(bk, bp, bv, p.cand, sk, sp);
    }
    grid.sync();
    int i = 0;
    double inv_prev = 1.0;                                          // reciprocal pivot of step i - 1
    for (; i < mn; ++i) {
        const long long pos = grid_winner(p.cand, G, &s_pos, &s_val);
        const int pr = (int)(pos % m), pc = (int)(pos / m);
        const double diag = s_val;                                  // the candidate carries the entry: no read-back, no barrier
        if (diag == 0.0) break;                                     // full_piv_lu.rs:73-76: the rest of the matrix is zero
        if (gtid == 0) { p.p_row[i] = pr; p.p_col[i] = pc; }
        // ---- phase S
        if (pc != i)
            for (long long r = gtid; r < m; r += nthreads) {
                if (r == i || r == pr) continue;
                const double t = a[r + i * lda]; a[r + i * lda] = a[r + pc * lda]; a[r + pc * lda] = t;
            }
        for (long long j = gtid; j < n; j += nthreads) {
            if (j == i || j == pc) continue;
            double t1 = a[i + j * lda], t2 = a[pr + j * lda];
            if (j == i - 1) { t1 = __dmul_rn(t1, inv_prev); t2 = __dmul_rn(t2, inv_prev); }      // multipliers of step i - 1
            if (pr != i) { a[i + j * lda] = t2; a[pr + j * lda] = t1; }
            else if (j == i - 1) a[i + j * lda] = t1;
        }
        if (gtid == nthreads - 1) {                                 // the corners: column swap, then row swap
            const double oii = a[i + i * lda], oip = a[i + pc * lda], opi = a[pr + i * lda], opp = a[pr + pc * lda];
            a[i + i * lda] = opp; a[i + pc * lda] = opi; a[pr + i * lda] = oip; a[pr + pc * lda] = oii;
        }
        if (i > 0)
            for (long long r = i + 1 + gtid; r < m; r += nthreads)
                if (r != pr) a[r + (i - 1) * lda] = __dmul_rn(a[r + (i - 1) * lda], inv_prev);
        grid.sync();
        // ---- phase U: RU rows per lane in flight
        const double inv_diag = 1.0 / diag;                         // lu.rs:345
        const double* ci = a + i * lda;                             // column i: unscaled until the next phase S
        double bk = -2.0, bv = 0.0; long long bp = 0x7fffffffffffffffll;
        for (int j = i + 1 + gwarp; j < n; j += nwarps) {
            double* cj = a + j * lda;
            const double npiv = -cj[i];                             // -pivot_row[k]   (lu.rs:353-356)
            for (int r0 = i + 1 + lane; r0 < m; r0 += 32 * RU) {
                double x[RU], y[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; y[u] = r < m ? cj[r] : 0.0; }
#pragma unroll
                for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; x[u] = r < m ? ci[r] : 0.0; }      // the pivot column: L1 hits after the first warp
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int r = r0 + 32 * u;
                    if (r < m) {
                        const double coeff = __dmul_rn(x[u], inv_diag);                  // coeffs *= inv_diag (the value the reference stores)
                        const double v = __dadd_rn(__dmul_rn(npiv, coeff), y[u]);        // axpy: a * x + y, never fused
                        cj[r] = v;
                        const double k = cand_key(v, r == i + 1 && j == i + 1);
                        const long long pos2 = r + (long long)j * m;
                        if (cand_better(k, pos2, bk, bp)) { bk = k; bp = pos2; bv = v; }
                    }
                }
            }
        }
        block_publish(bk, bp, bv, p.cand, sk, sp);
        inv_prev = inv_diag;
        grid.sync();
    }
    // the last pivot column's multipliers
    if (i > 0)
        for (long long r = i + gtid; r < m; r += nthreads) a[r + (i - 1) * lda] = __dmul_rn(a[r + (i - 1) * lda], inv_prev);
    if (gtid == 0) *p.steps_done = i;
    for (long long s = i + gtid; s < mn; s += nthreads) { p.p_row[s] = (int)s; p.p_col[s] = (int)s; }      // steps that never happened
}

// ------------------------------------------------------------------------------------------------------------------
// ColPivQR.  Two grid barriers per step:
//   phase H: CTA 0 turns the pivot column (still at its old place pc) into the unit axis of the reflection, written to
//            column i while the old column i moves to pc; the other CTAs swap the finished rows ..i of the two columns;
//   phase R: reflection of the columns right of i fused with the search for the next pivot.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sred) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) sred[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < pv::T / 32; ++w) t += sred[w];
    return t;
}

__global__ void __launch_bounds__(pv::T, 1) col_piv_qr_kernel(const PivotedParams p) {
    using namespace pv;
    cg::grid_group grid = cg::this_grid();
    __shared__ double sk[2 * (T / 32)];
    __shared__ long long sp[T / 32];
    __shared__ long long s_pos;
    __shared__ double s_val;
    __shared__ double sred[T / 32];
    __shared__ double s_dot[2][T / 32];
    constexpr int CR = 8192 / T;                                    // entries of a column a thread keeps (CTA-per-column sweep)
    constexpr int kColCtaMin = 2048;                                // shorter columns: warp per column (thresholds 512 .. 2048 measure the same)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int m = p.m, n = p.n, mn = min(m, n);
    const long long lda = p.lda;
    double* a = p.a;
    const int gwarp = cta * (T / 32) + warp, nwarps = G * (T / 32);
    const long long gtid = (long long)cta * T + tid, nthreads = (long long)G * T;
    {
        double bk = -2.0, bv = 0.0; long long bp = 0x7fffffffffffffffll;
        for (int j = gwarp; j < n; j += nwarps)
            for (int r = lane; r < m; r += 32) {
                const double x = a[r + j * lda];
                const double k = cand_key(x, r == 0 && j == 0);
                const long long pos = r + (long long)j * m;
                if (cand_better(k, pos, bk, bp)) { bk = k; bp = pos; bv = x; }
            }
        block_publish(bk, bp, bv, p.cand, sk, sp);
    }
    grid.sync();
    for (int i = 0; i < mn; ++i) {
        const long long pos = grid_winner(p.cand, G, &s_pos, &s_val);
        const int pc = (int)(pos / m);
        if (gtid == 0) p.p_col[i] = pc;
        double* hh = p.hh + 2 * (i & 1);
        // ---- phase H
        if (cta == 0) {
            // column pc, rows i.. -> unit axis (householder.rs:19-53) stored in column i; old column i -> column pc.
            // Up to 4096 / T entries per thread live in registers (one round of loads for the three passes); longer columns
            // are re-read.
            double* ci = a + i + i * lda;
            double* cp = a + i + pc * lda;
            const int len = m - i;
            constexpr int R = 4096 / T;                             // columns of up to 4096 entries stay in registers
            const bool in_regs = len <= R * T;
            double xr[R], oi[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int r = tid + u * T;
                xr[u] = (in_regs && r < len) ? cp[r] : 0.0;
                oi[u] = (in_regs && r < len && pc != i) ? ci[r] : 0.0;
            }
            double s = 0.0;
            if (in_regs) {
#pragma unroll
                for (int u = 0; u < R; ++u) s = fma(xr[u], xr[u], s);
            } else {
                for (int r = tid; r < len; r += T) { const double x = cp[r]; s = fma(x, x, s); }
            }
            const double sq = block_sum(s, sred);
            const double nrm = sqrt(sq);
            const double x0 = cp[0];
            const double modulus = x0 >= 0.0 ? x0 : -x0, sign = x0 >= 0.0 ? 1.0 : -1.0;      // simba to_exp
            const double signed_norm = sign * nrm;
            const double factor = (sq + modulus * nrm) * 2.0;
            __syncthreads();                                        // cp[0] has been read by everyone before it may change
            if (factor != 0.0) {
                const double sf = sqrt(factor);
                double s2 = 0.0;
                if (in_regs) {
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        const int r = tid + u * T;
                        xr[u] = r < len ? (r == 0 ? x0 + signed_norm : xr[u]) / sf : 0.0;    // unscale_mut
                        s2 = fma(xr[u], xr[u], s2);
                    }
                } else {
                    for (int r = tid; r < len; r += T) {
                        const double v = (r == 0 ? x0 + signed_norm : cp[r]) / sf;
                        s2 = fma(v, v, s2);
                    }
                }
                const double nn = sqrt(block_sum(s2, sred));        // normalize_mut
                if (in_regs) {
#pragma unroll
                    for (int u = 0; u < R; ++u) {
                        const int r = tid + u * T;
                        if (r < len) { ci[r] = xr[u] / nn; if (pc != i) cp[r] = oi[u]; }
                    }
                } else {
                    for (int r = tid; r < len; r += T) {
                        const double old_i = ci[r];
                        const double v = (r == 0 ? x0 + signed_norm : cp[r]) / sf;
                        ci[r] = v / nn;
                        if (pc != i) cp[r] = old_i;
                    }
                }
                if (tid == 0) { p.diag[i] = -signed_norm; hh[0] = signbit(-signed_norm) ? -1.0 : 1.0; hh[1] = 1.0; }   // signum of the returned norm; reflected
            } else {
                // an all-zero column is not reflected (householder.rs:36-48); it still changes places with column i
                for (int r = tid; r < len; r += T) {
                    const double old_i = ci[r];
                    ci[r] = r == 0 ? x0 + signed_norm : cp[r];
                    if (pc != i) cp[r] = old_i;
                }
                if (tid == 0) { p.diag[i] = signed_norm; hh[0] = 1.0; hh[1] = 0.0; }
            }
            if (G == 1 && pc != i)
                for (int r = tid; r < i; r += T) { const double t = a[r + i * lda]; a[r + i * lda] = a[r + pc * lda]; a[r + pc * lda] = t; }
        } else if (pc != i) {
            for (long long r = gtid - T; r < i; r += nthreads - T) {                      // rows ..i of the two columns
                const double t = a[r + i * lda]; a[r + i * lda] = a[r + pc * lda]; a[r + pc * lda] = t;
            }
        }
        grid.sync();
        // ---- phase R: reflect the columns right of i (rows i..) and search the next pivot (rows / columns i + 1..)
        const double sign = __ldcg(hh + 0);
        const bool reflected = __ldcg(hh + 1) != 0.0;
        const double* axis = a + i + i * lda;
        const int len = m - i;
        double bk = -2.0, bv = 0.0; long long bp = 0x7fffffffffffffffll;
        if (reflected && len > kColCtaMin && len <= CR * T) {
            // Long columns: a warp would read its column twice (dot product, then reflection) and the second read misses
            // the L1.  Here a CTA owns a column, every thread keeps CR entries of it in registers between the dot product
            // (one CTA-wide sum: warp shuffles + one barrier, double-buffered slots) and the reflection, and the next
            // column's loads are issued before the current one is reduced: one read + one write per element.
            double y[2][CR];
            int j = i + 1 + cta;
            if (j < n) {
#pragma unroll
                for (int u = 0; u < CR; ++u) { const int r = tid + u * T; y[0][u] = r < len ? a[i + r + (long long)j * lda] : 0.0; }
            }
            int par = 0;
            auto step = [&](double (&yc)[CR], double (&yn)[CR], int jc) {
                const int jn = jc + G;
                if (jn < n) {
#pragma unroll
                    for (int u = 0; u < CR; ++u) { const int r = tid + u * T; yn[u] = r < len ? a[i + r + (long long)jn * lda] : 0.0; }
                }
                double d = 0.0;
#pragma unroll
                for (int u = 0; u < CR; ++u) { const int r = tid + u * T; if (r < len) d = fma(axis[r], yc[u], d); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (lane == 0) s_dot[par][warp] = d;
                __syncthreads();
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < T / 32; ++w) tot += s_dot[par][w];
                par ^= 1;
                const double factor = tot * (sign * -2.0);           // reflection.rs:76-79, bias = 0
                double* cj = a + i + (long long)jc * lda;
#pragma unroll
                for (int u = 0; u < CR; ++u) {
                    const int r = tid + u * T;
                    if (r < len) {
                        const double v = __dadd_rn(__dmul_rn(factor, axis[r]), __dmul_rn(sign, yc[u]));   // axpy(factor, axis, sign)
                        cj[r] = v;
                        if (r >= 1) {
                            const double k = cand_key(v, r == 1 && jc == i + 1);
                            const long long pos2 = (i + r) + (long long)jc * m;
                            if (cand_better(k, pos2, bk, bp)) { bk = k; bp = pos2; bv = v; }
                        }
                    }
                }
            };
            for (; j < n; j += 2 * G) {
                step(y[0], y[1], j);
                if (j + G < n) step(y[1], y[0], j + G);
            }
        } else
        for (int j = i + 1 + gwarp; j < n; j += nwarps) {
            double* cj = a + i + j * lda;
            double factor = 0.0;
            if (reflected) {
                double d = 0.0;
                for (int r0 = lane; r0 < len; r0 += 32 * RU) {
                    double y[RU];
#pragma unroll
                    for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; y[u] = r < len ? cj[r] : 0.0; }
#pragma unroll
                    for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; if (r < len) d = fma(axis[r], y[u], d); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                factor = d * (sign * -2.0);                          // reflection.rs:76-79, bias = 0
            }
            for (int r0 = lane; r0 < len; r0 += 32 * RU) {
                double x[RU], y[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; y[u] = r < len ? cj[r] : 0.0; }
#pragma unroll
                for (int u = 0; u < RU; ++u) { const int r = r0 + 32 * u; x[u] = r < len ? axis[r] : 0.0; }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int r = r0 + 32 * u;
                    if (r < len) {
                        double v = y[u];
                        if (reflected) { v = __dadd_rn(__dmul_rn(factor, x[u]), __dmul_rn(sign, v)); cj[r] = v; }   // axpy(factor, axis, sign)
                        if (r >= 1) {
                            const double k = cand_key(v, r == 1 && j == i + 1);
                            const long long pos2 = (i + r) + (long long)j * m;
                            if (cand_better(k, pos2, bk, bp)) { bk = k; bp = pos2; bv = v; }
                        }
                    }
                }
            }
        }
        block_publish(bk, bp, bv, p.cand, sk, sp);
        grid.sync();
    }
}

static int pivoted_grid(const void* kernel, size_t m, size_t n) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, pv::T, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    // one warp per column and pass: more CTAs than columns / 8 only add barrier latency
    const size_t want = std::max<size_t>(1, ceil_div(std::max(n, (size_t)1), (size_t)(pv::T / 32)));
    (void)m;
    return (int)std::min<size_t>((size_t)ctx().sm_count, want);
}

// p_row / p_col: DEVICE int arrays of min(m, n) entries; steps (DEVICE int): elimination steps done.
int full_piv_lu_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, int* p_row, int* p_col, int* steps) {
    const size_t mn = std::min(m, n);
    if (mn == 0) return NA_OK;
    if (lda < m) { set_error("full_piv_lu: lda < m"); return NA_EINVAL; }
    if (m > 0x7fffff00ull || n > 0x7fffff00ull) { set_error("full_piv_lu: dimension exceeds 2^31"); return NA_EINVAL; }
    const int G = pivoted_grid((const void*)full_piv_lu_kernel, m, n);
    Scratch cand;
    NAB_TRY(cand.alloc((size_t)G * sizeof(pv::Cand), s));
    PivotedParams p{a, (long long)lda, (int)m, (int)n, p_row, p_col, nullptr, cand.as<pv::Cand>(), nullptr, steps};
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)full_piv_lu_kernel, dim3((unsigned)G), dim3(pv::T), args, 0, s));
    count_launch();
    return NA_OK;
}

// p_col: DEVICE int array of min(m, n) entries; diag: DEVICE, min(m, n) signed norms (nalgebra's ColPivQR::diag).
int col_piv_qr_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diag, int* p_col) {
    const size_t mn = std::min(m, n);
    if (mn == 0) return NA_OK;
    if (lda < m) { set_error("col_piv_qr: lda < m"); return NA_EINVAL; }
    if (m > 0x7fffff00ull || n > 0x7fffff00ull) { set_error("col_piv_qr: dimension exceeds 2^31"); return NA_EINVAL; }
    const int G = pivoted_grid((const void*)col_piv_qr_kernel, m, n);
    Scratch cand, hh;
    NAB_TRY(cand.alloc((size_t)G * sizeof(pv::Cand), s));
    NAB_TRY(hh.alloc(4 * sizeof(double), s));      // (sign, reflected) x step parity
    PivotedParams p{a, (long long)lda, (int)m, (int)n, nullptr, p_col, diag, cand.as<pv::Cand>(), hh.as<double>(), nullptr};
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)col_piv_qr_kernel, dim3((unsigned)G), dim3(pv::T), args, 0, s));
    count_launch();
    return NA_OK;
}

}  // namespace nab

using namespace nab;

// host side of both calls: permutation sequences as (i, i2) pairs with i != i2 only, like PermutationSequence::append_permutation
static size_t pairs_from_pivots(const std::vector<int>& piv, size_t count, size_t* swaps) {
    size_t len = 0;
    if (swaps) {
        for (size_t i = 0; i < count; ++i)
            if ((size_t)piv[i] != i) { swaps[2 * len] = i; swaps[2 * len + 1] = (size_t)piv[i]; ++len; }
        for (size_t i = 2 * len; i < 2 * piv.size(); ++i) swaps[i] = 0;
    }
    return len;
}

extern "C" {

int na_full_piv_lu_f64_dev(size_t m, size_t n, double* a, size_t lda, size_t* p_swaps, size_t* np, size_t* q_swaps, size_t* nq, void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t mn = std::min(m, n);
    if (np) *np = 0;
    if (nq) *nq = 0;
    if (mn == 0) return NA_OK;
    if (!a) { set_error("full_piv_lu: null matrix"); return NA_EINVAL; }
    Scratch piv;
    NAB_TRY(piv.alloc((2 * mn + 1) * sizeof(int), s));
    int* d = piv.as<int>();
    NAB_TRY(full_piv_lu_device(s, m, n, a, lda, d, d + mn, d + 2 * mn));
    std::vector<int> h(2 * mn + 1);
    NAB_CUDA(cudaMemcpyAsync(h.data(), d, (2 * mn + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    const size_t steps = (size_t)h[2 * mn];
    std::vector<int> pr(h.begin(), h.begin() + mn), pc(h.begin() + mn, h.begin() + 2 * mn);
    const size_t lp = pairs_from_pivots(pr, steps, p_swaps), lq = pairs_from_pivots(pc, steps, q_swaps);
    if (np) *np = lp;
    if (nq) *nq = lq;
    return NA_OK;
}

int na_full_piv_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* p_swaps, size_t* np, size_t* q_swaps, size_t* nq) {
    NAB_TRY(ensure_init());
    if (np) *np = 0;
    if (nq) *nq = 0;
    if (std::min(m, n) == 0) return NA_OK;
    if (!a || lda < m) { set_error("full_piv_lu: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d; size_t ldd;
    NAB_TRY(upload_matrix(s, d, ldd, a, lda, m, n));
    NAB_TRY(na_full_piv_lu_f64_dev(m, n, d.as<double>(), ldd, p_swaps, np, q_swaps, nq, s));
    NAB_TRY(download_matrix(s, a, lda, d.as<double>(), ldd, m, n));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

int na_col_piv_qr_f64_dev(size_t m, size_t n, double* a, size_t lda, double* diag, size_t* p_swaps, size_t* np, void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t mn = std::min(m, n);
    if (np) *np = 0;
    if (mn == 0) return NA_OK;
    if (!a || !diag) { set_error("col_piv_qr: null argument"); return NA_EINVAL; }
    Scratch piv;
    NAB_TRY(piv.alloc(mn * sizeof(int), s));
    NAB_TRY(col_piv_qr_device(s, m, n, a, lda, diag, piv.as<int>()));
    std::vector<int> h(mn);
    NAB_CUDA(cudaMemcpyAsync(h.data(), piv.p, mn * sizeof(int), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    const size_t lp = pairs_from_pivots(h, mn, p_swaps);
    if (np) *np = lp;
    return NA_OK;
}

int na_col_piv_qr_f64(size_t m, size_t n, double* a, size_t lda, double* diag, size_t* p_swaps, size_t* np) {
    NAB_TRY(ensure_init());
    const size_t mn = std::min(m, n);
    if (np) *np = 0;
    if (mn == 0) return NA_OK;
    if (!a || !diag || lda < m) { set_error("col_piv_qr: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d, dd; size_t ldd;
    NAB_TRY(upload_matrix(s, d, ldd, a, lda, m, n));
    NAB_TRY(dd.alloc(mn * sizeof(double), s));
    NAB_TRY(na_col_piv_qr_f64_dev(m, n, d.as<double>(), ldd, dd.as<double>(), p_swaps, np, s));
    NAB_TRY(download_matrix(s, a, lda, d.as<double>(), ldd, m, n));
    NAB_CUDA(cudaMemcpyAsync(diag, dd.p, mn * sizeof(double), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

}  // extern "C"
