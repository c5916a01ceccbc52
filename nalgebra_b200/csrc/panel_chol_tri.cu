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
// POTF2: in-place lower Cholesky of an n x n (n <= 128) diagonal block held in shared memory.
// Follows Cholesky::new_internal's pivot rule (/root/reference/src/linalg/cholesky.rs:237-268):
// pivot <= 0 or NaN -> use `sub` when allowed (and itself > 0), else record the failing column;
// the column is divided (true division) by sqrt(pivot).  The strict upper triangle is not touched.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
potf2_kernel(double* __restrict__ a, long long lda, int n, int use_sub, double sub, long long col0, unsigned long long* fail_col) {
    extern __shared__ double sm[];
    const int LDS = IB + 1;
    double* s = sm;                                   // n x n, ld = 129
    __shared__ double s_rdiag;
    __shared__ int s_fail;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx % n, j = idx / n;
        if (i >= j) s[i + j * LDS] = a[i + j * lda];
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
        if (tid == 0) {
            double d = s[j + j * LDS];
            bool ok = d > 0.0;                         // false for NaN and for <= 0
            if (!ok && use_sub && sub > 0.0) { d = sub; ok = true; }
            if (!ok) {
                atomicMin(fail_col, (unsigned long long)(col0 + j));
                s_fail = 1;
                d = 1.0;                               // keep going on garbage; the driver reports NOT_PD
            }
            const double sd = sqrt(d);
            s[j + j * LDS] = sd;
            s_rdiag = sd;
        }
        __syncthreads();
        const double sd = s_rdiag;
        for (int i = j + 1 + tid; i < n; i += nt) s[i + j * LDS] = s[i + j * LDS] / sd;
        __syncthreads();
        // trailing update of the lower triangle: a[i,k] -= a[i,j]*a[k,j], j < k <= i
        const int rem = n - j - 1;
        for (int idx = tid; idx < rem * rem; idx += nt) {
            const int i = j + 1 + idx % rem, k = j + 1 + idx / rem;
            if (i >= k) s[i + k * LDS] -= s[i + j * LDS] * s[k + j * LDS];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx % n, j = idx / n;
        if (i >= j) a[i + j * lda] = s[i + j * LDS];
    }
}

int potf2(cudaStream_t st, double* a, size_t lda, int n, int use_sub, double sub, size_t col0, unsigned long long* fail_col) {
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IB * (IB + 1) * 8); });
    potf2_kernel<<<1, 256, IB * (IB + 1) * 8, st>>>(a, (long long)lda, n, use_sub, sub, (long long)col0, fail_col);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// TRTRI of diagonal blocks.  M = op(T) is an n x n effectively lower (eff_lower=1) or upper
// triangular matrix given with element strides (rs, cs).  Block b covers rows/cols [b*128, ...).
// out[b] (128 x 128 column-major, ld 128) receives inverse(M_bb), zero outside the triangle and
// identity-padded when the last block is short.  unit: implicit unit diagonal.  diag_abs != null:
// the diagonal is |diag_abs[i]| instead of M[i,i] (nalgebra's QR keeps R's diagonal in `diag`,
// /root/reference/src/linalg/qr.rs:224-256).
//
// One CTA of 128 threads per block, thread j owns column j of the inverse.  A single 128x128
// shared buffer holds M's strict lower triangle and, transposed into the upper triangle + diagonal,
// the inverse being built (thread j touches row j only: conflict-free); upper-triangular blocks
// are handled by reversing the index order, which maps them to lower-triangular ones.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
trtri_blocks_kernel(const double* __restrict__ t, long long rs, long long cs, long long n, int eff_lower, int unit,
                    const double* __restrict__ diag_abs, double* __restrict__ out) {
    extern __shared__ double sm[];
    double* buf = sm;                  // 128 x 128, ld 128
    double* rdiag = sm + IB * IB;      // 128
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long base = (long long)b * IB;
    const int nb = (int)min((long long)IB, n - base);
    // local index li <-> global index: lower: base+li ; upper: base + (nb-1-li)
    auto gidx = [&](int li) -> long long { return eff_lower ? base + li : base + (nb - 1 - li); };
    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        const int i = idx % nb, k = idx / nb;
        if (i > k) buf[i + k * IB] = t[gidx(i) * rs + gidx(k) * cs];
    }
    if (tid < nb) {
        double d = 1.0;
        if (!unit) d = diag_abs ? fabs(diag_abs[gidx(tid)]) : t[gidx(tid) * rs + gidx(tid) * cs];
        rdiag[tid] = 1.0 / d;
    }
    __syncthreads();
    const int j = tid;
    if (j < nb) {
        // X[i,j] for i >= j, stored at buf[j + i*IB]
        buf[j + j * IB] = rdiag[j];
        for (int i = j + 1; i < nb; ++i) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int k = j;
            for (; k + 3 < i; k += 4) {
                s0 += buf[i + k * IB] * buf[j + k * IB];
                s1 += buf[i + (k + 1) * IB] * buf[j + (k + 1) * IB];
                s2 += buf[i + (k + 2) * IB] * buf[j + (k + 2) * IB];
                s3 += buf[i + (k + 3) * IB] * buf[j + (k + 3) * IB];
            }
            for (; k < i; ++k) s0 += buf[i + k * IB] * buf[j + k * IB];
            buf[j + i * IB] = -((s0 + s1) + (s2 + s3)) * rdiag[i];
        }
    }
    __syncthreads();
    // write out: out_b[gi, gj] (block-local global order) = X[li, lj]
    double* ob = out + (long long)b * IB * IB;
    for (int idx = tid; idx < IB * IB; idx += blockDim.x) {
        const int r = idx % IB, c = idx / IB;     // block-local position in global order
        double v = (r == c) ? 1.0 : 0.0;          // identity padding
        if (r < nb && c < nb) {
            const int li = eff_lower ? r : nb - 1 - r, lj = eff_lower ? c : nb - 1 - c;
            v = (li >= lj) ? buf[lj + li * IB] : 0.0;
        }
        ob[r + c * IB] = v;
    }
}

int trtri_blocks(cudaStream_t st, const double* t, ptrdiff_t rs, ptrdiff_t cs, size_t n, bool eff_lower, bool unit,
                 const double* diag_abs, double* out) {
    if (n == 0) return NA_OK;
    static std::once_flag once;
    const int smem = (IB * IB + IB) * 8;
    std::call_once(once, [smem] { cudaFuncSetAttribute(trtri_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    const int nblk = (int)ceil_div(n, IB);
    trtri_blocks_kernel<<<nblk, 128, smem, st>>>(t, rs, cs, (long long)n, eff_lower ? 1 : 0, unit ? 1 : 0, diag_abs, out);
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
