// dmma_stream.cuh -- C -= A * X for a tall C streamed once through the FP64 tensor pipe, one warp per row slice.
//
// The panel-internal updates of the blocked factorizations (LU: A22 -= A21 * U12 with K = leaf width <= 64; QR: C -= V * X
// with K = 32) have K so small that the 128 x 128 x 16 tile engine of dgemm.cu spends its time in prologues and
// epilogues.  They are bandwidth problems: every element of C is read and written once and takes K FMAs.  Here the
// warps of a CTA share a slab of rows in 8-row DMMA tiles (mma.sync m8n8k4), warp w taking tiles w, w + 8, ... -- at any
// moment the CTA works on 64 consecutive rows of the same columns, so DRAM pages, L2 lines and TLB entries are shared
// instead of every warp touching 64-byte pieces of its own lines.  The A fragments of a tile (8 rows x K) live in registers, X
// (K x nc) in shared memory, and C fragments come straight from global memory -- an accumulator fragment is 8 rows x 2
// adjacent columns per lane quad, i.e. whole 32-byte sectors -- 8 * NTB columns (2 * NTB loads per thread) at a time, with
// the next batch of columns (or the next row tile's first batch) in flight while the current ones are in the pipe.
#pragma once
#include "ptx.cuh"

namespace nab {

// NP: row stride of X in shared memory, = 4 (mod 16) doubles so that a B fragment (4 consecutive k x 8 consecutive
// columns) hits 16 distinct 8-byte banks; the array must be readable up to column 8 * NTB * ceil(ceil(nc / 8) / NTB) + 7
// of its last row (values there are never stored).
// c0: the slab's first row, column 0; nrows: rows of the slab; this warp takes the 8-row tiles tile0, tile0 + tstride, ...
// load_a(na, r): na[ks] = -A[r + lane / 4][4 * ks + lane % 4] for the row tile starting at slab row r (0 for rows
// outside the slab).  REVERSE walks the row tiles last to first.
template <int KS, int NP, bool REVERSE, int NTB = 8, class LoadA>
__device__ __forceinline__ void dmma_stream_update(double* c0, long long ldc, int nrows, int nc, const double* xs, int lane, int tile0, int tstride,
                                                   LoadA&& load_a) {
    const int g8 = lane >> 2, q4 = lane & 3;
    const int tiles = (nrows + 7) >> 3;
    const int ntr = tiles > tile0 ? (tiles - tile0 + tstride - 1) / tstride : 0, ntc = (nc + 7) >> 3, nbatch = (ntc + NTB - 1) / NTB;
    if (ntr == 0 || ntc == 0) return;
    double na[KS], na2[KS], cur[NTB][2], nxt[NTB][2];
    auto tile_row = [&](int t) { return 8 * (tile0 + tstride * (REVERSE ? ntr - 1 - t : t)); };
    auto load_c = [&](double (&dst)[NTB][2], int r, int b) {
        const bool rok = r + g8 < nrows;
        const double* crow = c0 + (r + g8) + (long long)(2 * q4) * ldc;
#pragma unroll
        for (int t = 0; t < NTB; ++t) {
            const int c = 8 * (NTB * b + t) + 2 * q4;
            dst[t][0] = (rok && c < nc) ? __ldcg(crow + (long long)(8 * (NTB * b + t)) * ldc) : 0.0;
            dst[t][1] = (rok && c + 1 < nc) ? __ldcg(crow + (long long)(8 * (NTB * b + t) + 1) * ldc) : 0.0;
        }
    };
    int t_r = 0, b = 0;
    load_a(na, tile_row(0));
    load_c(cur, tile_row(0), 0);
    const double* xq = xs + q4 * NP + g8;
    while (true) {
        int t_r2 = t_r, b2 = b + 1;
        if (b2 == nbatch) { b2 = 0; ++t_r2; }
        const bool more = t_r2 < ntr;
        if (more) {
            if (b2 == 0) load_a(na2, tile_row(t_r2));
            load_c(nxt, tile_row(t_r2), b2);
        }
        {
            // k-steps outermost: the eight column tiles are independent accumulator chains, so the tensor pipe never
            // waits for the result of the DMMA it has just been given
            const double* xb = xq + 8 * NTB * b;
            const int nt_b = ntc - NTB * b;                      // tiles of this batch that exist (>= NTB: all)
            if (nt_b >= NTB) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double xf[NTB];
#pragma unroll
                    for (int t = 0; t < NTB; ++t) xf[t] = xb[(4 * ks) * NP + 8 * t];
#pragma unroll
                    for (int t = 0; t < NTB; ++t) ptx::dmma884(cur[t][0], cur[t][1], na[ks], xf[t]);
                }
            } else {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double xf[NTB];
#pragma unroll
                    for (int t = 0; t < NTB; ++t) xf[t] = xb[(4 * ks) * NP + 8 * t];
#pragma unroll
                    for (int t = 0; t < NTB; ++t)
                        if (t < nt_b) ptx::dmma884(cur[t][0], cur[t][1], na[ks], xf[t]);
                }
            }
        }
        {
            const int r = tile_row(t_r);
            const bool rok = r + g8 < nrows;
            double* crow = c0 + (r + g8) + (long long)(2 * q4) * ldc;
#pragma unroll
            for (int t = 0; t < NTB; ++t) {
                const int c = 8 * (NTB * b + t) + 2 * q4;
                if (rok && c < nc) __stcg(crow + (long long)(8 * (NTB * b + t)) * ldc, cur[t][0]);
                if (rok && c + 1 < nc) __stcg(crow + (long long)(8 * (NTB * b + t) + 1) * ldc, cur[t][1]);
            }
        }
        if (!more) break;
#pragma unroll
        for (int t = 0; t < NTB; ++t) { cur[t][0] = nxt[t][0]; cur[t][1] = nxt[t][1]; }
        if (b2 == 0) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) na[ks] = na2[ks];
        }
        t_r = t_r2; b = b2;
    }
}

}  // namespace nab
