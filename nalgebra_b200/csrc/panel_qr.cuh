// panel_qr.cuh -- what the two GEQR2 leaf kernels (shared-memory resident rows: panel_qr.cu; register resident rows:
// panel_qr_reg.cu) share: the launch parameters, the self-validating 16-byte exchange words and the warp butterfly.
#pragma once
#include "common.cuh"

namespace nab {

struct Geqr2Params {
    double* a; long long lda;
    int m, w, rp;
    double* tau;               // [w] out
    double2* xch;              // [2][G][32]  (partial sum, seq) pairs, slot j = column j (slot c = sigma)
    double2* rowc;             // [2][32]     (a[c, j], seq) pairs published by the owner of row c
    double2* totx;             // [2][32]     (T_j, seq) totals published by the reducer CTA of slot j (two-stage exchange)
    int seq0;                  // sequence numbers already consumed in this workspace
};

__device__ __forceinline__ void st_pair(double2* p, double v, double seq) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(seq) : "memory");
}
__device__ __forceinline__ void ld_pair_raw(const double2* p, double& x, double& y) {
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
}

// vals[0..31] per lane -> returns in every lane l the sum over the warp of vals[l] (butterfly
// reduce-scatter: 31 shuffles instead of 160)
__device__ __forceinline__ double warp_reduce_scatter32(double (&vals)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool hi = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const double send = hi ? vals[i] : vals[i + step];
            const double keep = hi ? vals[i + step] : vals[i];
            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return vals[0];
}

constexpr size_t kGeqr2MaxCtas = 160;

}  // namespace nab
