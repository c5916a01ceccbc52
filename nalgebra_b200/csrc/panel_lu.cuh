// panel_lu.cuh -- device helpers shared by the LU panel kernels (panel_lu.cu, panel_lu_reg.cu):
// the self-validating (value, sequence) exchange words and the icamax candidate ordering
// (/root/reference/src/base/min_max.rs:221-240).
#pragma once
#include "common.cuh"

namespace nab {

// (value, seq) travel in one 16-byte word: a reader that sees the expected seq has the value.
__device__ __forceinline__ void lu_st_pair(double2* p, double v, double seq) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v), "d"(seq) : "memory");
}
__device__ __forceinline__ double lu_ld_pair(const double2* p, double seq) {
    double x, y;
    do {
        asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
    } while (y != seq);
    return x;
}

__device__ __forceinline__ void lu_ld_pair_raw(const double2* p, double& x, double& y) {
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p) : "memory");
}
// candidate header: (|value|, seq << 32 | row) in one 16-byte word
__device__ __forceinline__ double pack_seq_row(int seq, int row) {
    return __longlong_as_double(((long long)seq << 32) | (unsigned int)row);
}

// candidate ordering of icamax: larger value wins; on equal values the lower row wins.
__device__ __forceinline__ bool cand_better(double v1, int r1, double v2, int r2) {
    return (v1 > v2) || (v1 == v2 && r1 < r2);
}
// |x| as a pivot key: a NaN wins only at index 0 of the searched range (min_max.rs:221-240)
__device__ __forceinline__ double pivot_key(double x, bool first) {
    const double v = fabs(x);
    return (v != v) ? (first ? __longlong_as_double(0x7ff0000000000000LL) : -1.0) : v;
}

// Warp-wide winner of (v, r) under cand_better -- larger key, lowest row on ties -- with redux.sync on the
// key's bit pattern instead of five rounds of three shuffles (keys are |x| >= 0, +inf, or the markers -1 / -2,
// so `bits + 2` / 1 / 0 is monotone).  Every lane returns the winner; `tag` follows it (e.g. the CTA index).
__device__ __forceinline__ void warp_best(double& v, int& r, int& tag) {
    const unsigned FULL = 0xffffffffu;
    const unsigned long long kb = v >= 0.0 ? (unsigned long long)__double_as_longlong(v) + 2ull : (v == -1.0 ? 1ull : 0ull);
    const unsigned hi = (unsigned)(kb >> 32), lo = (unsigned)kb;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
    const bool is_max = hi == mhi && lo == mlo;
    const int mr = __reduce_min_sync(FULL, is_max ? r : 0x7fffffff);
    tag = __reduce_min_sync(FULL, (is_max && r == mr) ? tag : 0x7fffffff);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    v = mk >= 2ull ? __longlong_as_double((long long)(mk - 2ull)) : (mk == 1ull ? -1.0 : -2.0);
    r = mr;
}

}  // namespace nab
