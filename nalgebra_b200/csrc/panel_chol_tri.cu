// panel_chol_tri.cu -- single-CTA panel kernels: POTF2 (128-block Cholesky), TRTRI of the
// 128x128 diagonal blocks of a triangular matrix, plus small utility kernels.
//
// These are the latency-bound leaves of the blocked factorizations; everything O(n^3) goes through
// the DGEMM tile engine (dgemm.cu).
#include "common.cuh"
#include "kernels.cuh"

namespace nab {

constexpr int IB = kInvBlock;   // 128

// ------------------------------------------------------------------------------------------------
// Register-blocked 128x128 leaf kernels.  256 threads form a 16x16 grid; thread (tx, ty) owns the
// elements (i = tx + 16a, k = ty + 16b), a,b < 8, of the block in registers (cyclic distribution,
// so the shrinking active region stays balanced).  Each column/row step is one rank-1 update:
// the step's vector is broadcast through shared memory, 16 LDS + 64 predicated DFMA per thread,
// one __syncthreads per step.
// ------------------------------------------------------------------------------------------------
struct Tile16 {
    double r[8][8];
};

// X <- inverse of the n x n lower-triangular matrix held in sL (ld 128, diagonal included unless
// unit).  Forward substitution on the identity, right-looking: for step j, row j of X is scaled by
// 1/L[j,j] and eliminated from the rows below.  Result left in `x` (register tile).
__device__ __forceinline__ void tile_trtri_lower(const double* __restrict__ sL, double* __restrict__ rowbuf /*[3][128]*/,
                                                 int n, bool unit, Tile16& x, int tx, int ty) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) x.r[a][b] = (tx + 16 * a == ty + 16 * b) ? 1.0 : 0.0;
    double* rdiag = rowbuf + 256;                   // reciprocal diagonal, computed once, in parallel
    {
        const int t = tx + 16 * ty;
        if (t < n) rdiag[t] = unit ? 1.0 : 1.0 / sL[t + t * 128];
    }
    __syncthreads();
#pragma unroll
    for (int blk = 0; blk < 8; ++blk) {             // blk = j / 16 is a compile-time constant: static register indexing
        for (int jj = 0; jj < 16; ++jj) {
            const int j = blk * 16 + jj;
            if (j >= n) break;
            double* rb = rowbuf + (j & 1) * 128;
            if (tx == jj) {                             // owners of row j of X
                const double rd = rdiag[j];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const double v = x.r[blk][b] * rd;
                    x.r[blk][b] = v;
                    rb[ty + 16 * b] = v;
                }
            }
            __syncthreads();
            // only the 16-blocks that can change: rows i > j live in a >= blk, columns k <= j in b <= blk
            // ((8 - blk)(blk + 1) block products instead of 64: 15 on average)
            double lcol[8], xrow[8];
#pragma unroll
            for (int a = blk; a < 8; ++a) { const int i = tx + 16 * a; lcol[a] = (i > j && i < n) ? sL[i + j * 128] : 0.0; }
#pragma unroll
            for (int b = 0; b <= blk; ++b) { const int k = ty + 16 * b; xrow[b] = (k <= j) ? rb[k] : 0.0; }
#pragma unroll
            for (int a = blk; a < 8; ++a)
#pragma unroll
                for (int b = 0; b <= blk; ++b) x.r[a][b] -= lcol[a] * xrow[b];
        }
    }
}

// POTF2 + TRTRI of one diagonal block (n <= 128): in-place lower Cholesky following
// Cholesky::new_internal's pivot rule (/root/reference/src/linalg/cholesky.rs:237-268): pivot <= 0
// or NaN -> `sub` when allowed (and itself > 0), else the failing column is recorded; the column is
// divided (true division) by sqrt(pivot).  The strict upper triangle is neither read nor written.
// inv_out (128x128, ld 128) receives inverse(L), identity padded.
__global__ void __launch_bounds__(256, 1)
potf2_trtri_kernel(double* __restrict__ a, long long lda, int n, int use_sub, double sub, long long col0,
                   unsigned long long* fail_col, double* __restrict__ inv_out) {
    extern __shared__ double sm[];
    double* sL = sm;                   // 128 x 128 factor, column j filled at step j
    double* rowbuf = sm + 128 * 128;   // [2][128]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    Tile16 t;
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const int i = tx + 16 * aa, k = ty + 16 * b;
            t.r[aa][b] = (i < n && k < n && i >= k) ? a[i + k * lda] : 0.0;
        }
#pragma unroll
    for (int blk = 0; blk < 8; ++blk) {             // blk = j / 16 is a compile-time constant: static register indexing
        for (int jj = 0; jj < 16; ++jj) {
            const int j = blk * 16 + jj;
            if (j >= n) break;
            if (ty == jj) {                             // owners of column j: half a warp (lanes 16*(ty&1)..+15)
                // diagonal element lives in lane tx == jj, slot [blk][blk]
                double d = __shfl_sync(0xffffu << (tid & 16), t.r[blk][blk], jj + (tid & 16), 32);
                bool ok = d > 0.0;                      // false for NaN and for <= 0
                if (!ok && use_sub && sub > 0.0) { d = sub; ok = true; }
                if (!ok) {
                    if (tx == 0) atomicMin(fail_col, (unsigned long long)(col0 + j));
                    d = 1.0;                            // keep going on garbage; the driver reports NOT_PD
                }
                // diagonal = sqrt(pivot) exactly; the column is scaled by the reciprocal (the reference
                // divides by sqrt(pivot): same to ~1 ulp, without 8 serial FP64 divisions per thread)
                const double sd = sqrt(d);
                const double rsd = 1.0 / sd;
#pragma unroll
                for (int aa = 0; aa < 8; ++aa) {
                    const int i = tx + 16 * aa;
                    const double v = (i == j) ? sd : (i > j ? t.r[aa][blk] * rsd : 0.0);
                    sL[i + j * 128] = v;
                    t.r[aa][blk] = v;
                }
            }
            __syncthreads();
            // only the 16-blocks that can change and are stored: rows and columns > j live in blocks >= blk, the lower
            // triangle in aa >= b ((8 - blk)(9 - blk)/2 block products instead of 64: 15 on average)
            double ci[8], ck[8];
#pragma unroll
            for (int aa = blk; aa < 8; ++aa) { const int i = tx + 16 * aa; ci[aa] = (i > j) ? sL[i + j * 128] : 0.0; }
#pragma unroll
            for (int b = blk; b < 8; ++b) { const int k = ty + 16 * b; ck[b] = (k > j) ? sL[k + j * 128] : 0.0; }
#pragma unroll
            for (int aa = blk; aa < 8; ++aa)
#pragma unroll
                for (int b = blk; b <= aa; ++b) t.r[aa][b] -= ci[aa] * ck[b];   // entries with i < k are never stored
        }
    }
    // store L (lower triangle only)
#pragma unroll
    for (int b = 0; b < 8; ++b)
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const int i = tx + 16 * aa, k = ty + 16 * b;
            if (i < n && k < n && i >= k) a[i + k * lda] = t.r[aa][b];
        }
    __syncthreads();
    if (inv_out) {
        Tile16 x;
        tile_trtri_lower(sL, rowbuf, n, false, x, tx, ty);
#pragma unroll
        for (int b = 0; b < 8; ++b)
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const int i = tx + 16 * aa, k = ty + 16 * b;
                inv_out[i + k * 128] = (i < n && k < n) ? (i >= k ? x.r[aa][b] : 0.0) : (i == k ? 1.0 : 0.0);
            }
    }
}

int potf2(cudaStream_t st, double* a, size_t lda, int n, int use_sub, double sub, size_t col0, unsigned long long* fail_col,
          double* inv_out) {
    static std::once_flag once;
    const int smem = (128 * 128 + 384) * 8;
    std::call_once(once, [smem] { cudaFuncSetAttribute(potf2_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    potf2_trtri_kernel<<<1, 256, smem, st>>>(a, (long long)lda, n, use_sub, sub, (long long)col0, fail_col, inv_out);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// TRTRI of diagonal blocks.  M = op(T) is an n x n effectively lower (eff_lower=1) or upper
// triangular matrix given with element strides (rs, cs).  Block b covers rows/cols [b*128, ...).
// out[b] (128 x 128 column-major, ld 128) receives inverse(M_bb), zero outside the triangle and
// identity-padded when the last block is short.  unit: implicit unit diagonal.  diag_abs != null:
// the diagonal is |diag_abs[i]| instead of M[i,i] (nalgebra's QR keeps R's diagonal in `diag`,
// /root/reference/src/linalg/qr.rs:224-256).  Upper-triangular blocks are handled by reversing the
// index order, which maps them to lower-triangular ones.  One CTA per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
trtri_blocks_kernel(const double* __restrict__ t, long long rs, long long cs, long long n, int eff_lower, int unit,
                    const double* __restrict__ diag_abs, double* __restrict__ out) {
    extern __shared__ double sm[];
    double* sL = sm;
    double* rowbuf = sm + 128 * 128;
    const int b = blockIdx.x, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long base = (long long)b * IB;
    const int nb = (int)min((long long)IB, n - base);
    auto gidx = [&](int li) -> long long { return eff_lower ? base + li : base + (nb - 1 - li); };
    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int i = idx % nb, k = idx / nb;
        if (i > k) sL[i + k * 128] = t[gidx(i) * rs + gidx(k) * cs];
        else if (i == k) sL[i + k * 128] = unit ? 1.0 : (diag_abs ? fabs(diag_abs[gidx(i)]) : t[gidx(i) * rs + gidx(i) * cs]);
    }
    __syncthreads();
    Tile16 x;
    tile_trtri_lower(sL, rowbuf, nb, unit != 0, x, tx, ty);
    double* ob = out + (long long)b * IB * IB;
#pragma unroll
    for (int bb = 0; bb < 8; ++bb)
#pragma unroll
        for (int aa = 0; aa < 8; ++aa) {
            const int li = tx + 16 * aa, lj = ty + 16 * bb;          // local (lower) indices
            if (li < nb && lj < nb) {
                const int r = eff_lower ? li : nb - 1 - li, c = eff_lower ? lj : nb - 1 - lj;
                ob[r + c * IB] = (li >= lj) ? x.r[aa][bb] : 0.0;
            } else {
                ob[li + lj * IB] = (li == lj) ? 1.0 : 0.0;            // identity padding
            }
        }
}

int trtri_blocks(cudaStream_t st, const double* t, ptrdiff_t rs, ptrdiff_t cs, size_t n, bool eff_lower, bool unit,
                 const double* diag_abs, double* out) {
    if (n == 0) return NA_OK;
    static std::once_flag once;
    const int smem = (IB * IB + 384) * 8;
    std::call_once(once, [smem] { cudaFuncSetAttribute(trtri_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    const int nblk = (int)ceil_div(n, IB);
    trtri_blocks_kernel<<<nblk, 256, smem, st>>>(t, rs, cs, (long long)n, eff_lower ? 1 : 0, unit ? 1 : 0, diag_abs, out);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
// flag <- 1 if any diagonal entry (|diag_abs[i]| or t[i,i]) is exactly zero
// (solve_upper_triangular_mut returns false, /root/reference/src/linalg/solve.rs:169-171).
__global__ void zero_diag_check_kernel(const double* __restrict__ t, long long ldt, const double* __restrict__ diag_abs,
                                       long long n, int* __restrict__ flag) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = diag_abs ? diag_abs[i] : t[i + i * ldt];
        if (d == 0.0) *flag = 1;
    }
}
int zero_diag_check(cudaStream_t st, const double* t, size_t ldt, const double* diag_abs, size_t n, int* flag) {
    if (n == 0) return NA_OK;
    zero_diag_check_kernel<<<(int)std::min<size_t>(ceil_div(n, 256), 64), 256, 0, st>>>(t, (long long)ldt, diag_abs, (long long)n, flag);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// a(i,j) = (i == j) ? 1 : 0
__global__ void set_identity_kernel(double* __restrict__ a, long long lda, long long rows, long long cols) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % rows, c = idx / rows;
        a[r + c * lda] = (r == c) ? 1.0 : 0.0;
    }
}
int set_identity(cudaStream_t st, double* a, size_t lda, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return NA_OK;
    set_identity_kernel<<<(int)std::min<size_t>(ceil_div(rows * cols, 256), (size_t)ctx().sm_count * 8), 256, 0, st>>>(
        a, (long long)lda, (long long)rows, (long long)cols);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab
