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

namespace nab {

// ------------------------------------------------------------------------------------------------
// GEQR2: m x w panel (w <= 64), rows distributed over G co-resident CTAs, resident in shared
// memory; two grid-wide exchanges per column (|x|^2 and alpha; then the w-c-1 dot products v^T a_j).
//
// The exchange is self-validating: every published double travels with a sequence number in the
// same 16-byte word, so there is no separate barrier, fence or counter: readers poll the words they
// need until the sequence number matches.  Two buffers (by exchange parity) are enough because a
// CTA cannot publish exchange k+2 before every CTA has published k+1, i.e. finished reading k.
// ------------------------------------------------------------------------------------------------
struct Geqr2Params {
    double* a; long long lda;
    int m, w, rp;
    double* tau;               // [w] out
    double2* xch;              // [2][G][P] (value, seq) pairs, P = w + 1
    int seq0;                  // sequence numbers already consumed in this buffer (buffers are reused across panels)
};

__device__ __forceinline__ void st_pair(double2* p, double v, double seq) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(seq) : "memory");
}
__device__ __forceinline__ double2 ld_pair(const double2* p) {
    double2 r;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}

// All CTAs publish `np` partial values (part[0..np)) and receive the sums over CTAs in tot[0..np)
// (identical in every CTA: summed in CTA order).  stage: smem [G*np].
__device__ __forceinline__ void exchange_sum(const Geqr2Params& p, int G, int cta, int P, int np, double seq, int buf,
                                             const double* part, double* tot, double* stage) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double2* base = p.xch + (size_t)buf * G * P;
    for (int i = tid; i < np; i += nt) st_pair(base + (size_t)cta * P + i, part[i], seq);
    for (int idx = tid; idx < G * np; idx += nt) {
        const int g = idx / np, i = idx - g * np;
        const double2* src = base + (size_t)g * P + i;
        double2 v = ld_pair(src);
        while (v.y != seq) v = ld_pair(src);
        stage[idx] = v.x;
    }
    __syncthreads();
    for (int i = tid; i < np; i += nt) {
        double s = 0.0;
        for (int g = 0; g < G; ++g) s += stage[g * np + i];
        tot[i] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256, 1) geqr2_coop_kernel(const Geqr2Params p) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int w = p.w, rp = p.rp, P = w + 1;
    const int r_begin = cta * rp;
    const int nrows = max(0, min(rp, p.m - r_begin));
    double* s = sm;                          // [w][rp]
    double* part = s + (size_t)w * rp;       // [P]
    double* tot = part + P;                  // [P]
    double* wred = tot + P;                  // [8][P] per-warp partials
    double* stage = wred + 8 * P;            // [G*P]

    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) s[r + c * rp] = p.a[(long long)(r_begin + r) + (long long)c * p.lda];
    __syncthreads();

    const int ncol = min(w, p.m);
    int seq = p.seq0;
    for (int c = 0; c < ncol; ++c) {
        // ---- exchange 1: |x|^2 over rows >= c, and alpha = a[c,c] ----
        double ss = 0.0;
        for (int r = tid; r < nrows; r += nt)
            if (r_begin + r >= c) { const double v = s[r + c * rp]; ss += v * v; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) wred[warp] = ss;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < nt / 32; ++i) t += wred[i];
            part[0] = t;
            part[1] = (c >= r_begin && c < r_begin + nrows) ? s[(c - r_begin) + c * rp] : 0.0;
        }
        __syncthreads();
        ++seq;
        exchange_sum(p, G, cta, P, 2, (double)seq, seq & 1, part, tot, stage);
        const double sq = tot[0], alpha = tot[1];
        const double nrm = sqrt(sq);
        double beta = 0.0, tau = 0.0, scale = 0.0;
        if (nrm != 0.0) {                         // householder.rs:36: only an all-zero column is skipped
            beta = (alpha >= 0.0) ? -nrm : nrm;   // -sign(alpha)*|x|, sign(0) = +1 like simba's to_exp
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        if (cta == 0 && tid == 0) p.tau[c] = tau;
        // v = x * scale below the diagonal; the diagonal slot receives beta
        for (int r = tid; r < nrows; r += nt) {
            const int gr = r_begin + r;
            if (gr > c) s[r + c * rp] *= scale;
            else if (gr == c) s[r + c * rp] = beta;
        }
        __syncthreads();
        const int nrem = w - c - 1;
        if (nrem == 0 || tau == 0.0) continue;    // uniform across CTAs
        // ---- exchange 2: dots[j] = v^T a_j over rows >= c (v[c] = 1) ----
        for (int j0 = 0; j0 < nrem; j0 += 8) {    // 8 columns at a time in registers
            double acc[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0;
            for (int r = tid; r < nrows; r += nt) {
                const int gr = r_begin + r;
                if (gr < c) continue;
                const double v = (gr == c) ? 1.0 : s[r + c * rp];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    if (j0 + jj < nrem) acc[jj] += v * s[r + (c + 1 + j0 + jj) * rp];
            }
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[jj] += __shfl_xor_sync(0xffffffffu, acc[jj], o);
                if (lane == 0 && j0 + jj < nrem) wred[warp * P + j0 + jj] = acc[jj];
            }
        }
        __syncthreads();
        for (int j = tid; j < nrem; j += nt) {
            double t = 0.0;
            for (int i = 0; i < nt / 32; ++i) t += wred[i * P + j];
            part[j] = t;
        }
        __syncthreads();
        ++seq;
        exchange_sum(p, G, cta, P, nrem, (double)seq, seq & 1, part, tot, stage);
        // ---- a_j -= tau * dots[j] * v ----
        for (int r = tid; r < nrows; r += nt) {
            const int gr = r_begin + r;
            if (gr < c) continue;
            const double tv = tau * ((gr == c) ? 1.0 : s[r + c * rp]);
            for (int j = 0; j < nrem; ++j) s[r + (c + 1 + j) * rp] -= tv * tot[j];
        }
        __syncthreads();
    }
    for (int c = 0; c < w; ++c)
        for (int r = tid; r < nrows; r += nt) p.a[(long long)(r_begin + r) + (long long)c * p.lda] = s[r + c * rp];
}

constexpr size_t kGeqr2MaxCtas = 160;
size_t geqr2_workspace_bytes() { return 2 * kGeqr2MaxCtas * (kQrLeaf + 1) * sizeof(double2) + sizeof(int) * 4; }

// *seq_state (host) carries the sequence numbers consumed so far in this workspace.
int geqr2_panel(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, double* tau, void* ws, int* seq_state) {
    if (m == 0 || w == 0) return NA_OK;
    if (w > (size_t)kQrLeaf) { set_error("geqr2: panel too wide"); return NA_EINVAL; }
    const int sms = ctx().sm_count;
    const size_t P = w + 1;
    // smem: w*rp + 2P + 8P + G*P doubles
    size_t G = 1, rp = 0;
    for (;; ++G) {
        if (G > (size_t)std::min<int>(sms, (int)kGeqr2MaxCtas)) { set_error("geqr2: %zu x %zu panel does not fit in shared memory", m, w); return NA_EINVAL; }
        rp = round_up(ceil_div(m, G), 32);
        const size_t bytes = (w * rp + 10 * P + G * P) * sizeof(double);
        if (bytes <= 200 * 1024 && (rp <= 1024 || G * 2 > (size_t)sms)) break;   // prefer <= 1024 rows per CTA when SMs allow
    }
    G = ceil_div(m, rp);
    const size_t smem = (w * rp + 10 * P + G * P) * sizeof(double);
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(geqr2_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); });
    Geqr2Params p;
    p.a = a_panel; p.lda = (long long)lda; p.m = (int)m; p.w = (int)w; p.rp = (int)rp; p.tau = tau;
    p.xch = static_cast<double2*>(ws);
    p.seq0 = *seq_state;
    *seq_state += 2 * (int)w + 2;            // upper bound of the sequence numbers this panel uses (parity preserved)
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
__global__ void qr_signs_from_beta_kernel(const double* __restrict__ a, long long lda, const double* __restrict__ tau, int k,
                                          double* __restrict__ csign, double* __restrict__ diag) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double c = 1.0;
        csign[0] = c;
        for (int i = 0; i < k; ++i) {
            const double beta = a[i + (long long)i * lda];
            const bool refl = tau[i] != 0.0;
            diag[i] = refl ? c * beta : 0.0;
            if (refl) c = (beta < 0.0) ? -1.0 : 1.0;
            csign[i + 1] = c;
        }
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
__global__ void qr_convert_kernel(double* __restrict__ a, long long lda, long long m, long long n, int k,
                                  const double* __restrict__ tau, const double* __restrict__ csign, const double* __restrict__ diag) {
    const long long total = m * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx % m, j = idx / m;
        double* e = a + i + j * lda;
        if (i >= j) {                      // axis part of column j (j < k guaranteed since i < m, j <= i)
            if (j >= k) continue;
            const double t = tau[j];
            if (t == 0.0) { *e = 0.0; continue; }
            const double f = ((diag[j] < 0.0) ? 1.0 : -1.0) * sqrt(0.5 * t);      // -sign(diag_j)*sqrt(tau/2)
            *e = (i == j) ? f : f * *e;
        } else {                           // strict upper: row i < k
            *e = csign[i + 1] * *e;
        }
    }
}
int qr_convert_to_nalgebra(cudaStream_t st, double* a, size_t lda, size_t m, size_t n, const double* tau, double* csign, double* diag) {
    const size_t k = std::min(m, n);
    if (k == 0) return NA_OK;
    qr_signs_from_beta_kernel<<<1, 32, 0, st>>>(a, (long long)lda, tau, (int)k, csign, diag);
    NAB_LAUNCH_CHECK();
    qr_convert_kernel<<<(int)std::min<size_t>(ceil_div(m * n, 256), (size_t)ctx().sm_count * 16), 256, 0, st>>>(
        a, (long long)lda, (long long)m, (long long)n, (int)k, tau, csign, diag);
    NAB_LAUNCH_CHECK();
    return NA_OK;
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
