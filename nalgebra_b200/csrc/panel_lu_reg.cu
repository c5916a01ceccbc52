// panel_lu_reg.cu -- register-resident GETF2: the LU panel leaf with partial pivoting whose rows live
// in REGISTERS (one thread owns RPT whole rows of the W-column leaf, W * RPT = 64 doubles) and are never
// moved: pivoting is implicit.
//
// Reference semantics, column step of LU::new (/root/reference/src/linalg/lu.rs:103-119):
//   piv = icamax(A[i.., i]) + i  -- largest |x|, LOWEST index wins ties, NaN only wins at index 0
//                                   (src/base/min_max.rs:221-240)
//   diag == 0 -> skip the column; else swap whole rows, multiply the column by 1/diag (a reciprocal,
//   lu.rs:344-349) and apply the rank-1 update as unfused multiply + add (gauss_step, lu.rs:353-356).
//
// Why registers.  The shared-memory GETF2 (panel_lu.cu) costs ~6400 cycles per column whatever the panel
// height: its rank-1 update is shared-memory-bandwidth bound (LDS + STS of rows x (w - c) doubles per
// column) and it crosses 6-7 block barriers per column.  Here
//   * the rank-1 update is 2 (w - c) DP instructions per row on registers, its multiplier is the
//     thread's own register and the pivot row is a shared-memory broadcast;
//   * the pivot search is a redux.sync per warp + one block barrier;
//   * rows never move: every row carries its current LOGICAL position `pos` (what icamax's "lowest index
//     wins" refers to).  When the winner sits at position piv, the row at position c takes position piv
//     and the winner takes position c and retires (it then holds row c of U).  A retired row is simply
//     left alone; at the end every thread stores its rows at their final positions, which is exactly the
//     result of the reference's whole-row swaps inside the panel;
//   * the column loop is fully unrolled (all register indices are static).
// Across CTAs (G > 1, cooperative launch) the exchange per column is the self-validating scheme of
// panel_lu.cu: every CTA publishes its candidate header (|value|, seq << 32 | pos) and then the candidate's
// row entries as 16-byte (value, seq) words; every CTA polls the G headers, reduces them identically and
// reads the winner's row.  No fence, no counter, no grid barrier.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"
#include "panel_lu.cuh"

namespace nab {

constexpr int kRegThreads = 256;
constexpr int kNoRow = 0x7fffffff;

struct Getf2RegParams {
    double* a; long long lda;      // panel origin = A[j0, j0]
    int m, w;                      // panel rows / columns (w <= W)
    int rows_cta;                  // rows per CTA (<= 256 * RPT)
    int j0;                        // global row/col offset of the panel (for ipiv values)
    int* ipiv;                     // ipiv[j0 + c] = global pivot row of column c
    double2* xch;                  // [2][G][slot]: 32-byte header + 32-byte words of three row entries, by column parity
    int seq0;                      // sequence numbers already consumed in this workspace
    int* list;                     // optional: [0] = count (zeroed by the host), then dest[kRegListMax], src[kRegListMax]:
                                   // the rows this leaf moved, as global row indices ("row dest <- old row src")
};

// Row states kept in pos[]: >= 0 live at that logical position; kInvalidRow: padding beyond the panel;
// <= -2: retired as the pivot row of column (-2 - pos).
constexpr int kInvalidRow = -1;

// per-phase cycle counters of CTA 0 / thread 0 (tools/lu_reg_prof.py); off in the product build
__device__ long long g_getf2_reg_prof[16];
#ifdef NAB_GETF2_PROF
#define RPROF(i) do { const long long t_ = clock64(); t_acc[i] += t_ - t_prev; t_prev = t_; } while (0)
#else
#define RPROF(i) do { } while (0)
#endif
constexpr int kRegBlock = 8;          // columns per code block: the register window shrinks by this much per block

// 32-byte candidate header {x0, seq << 32 | pos, x1, (double)seq}: the candidate row's entries in the column being
// searched (x0, signed) and in the next one (x1).  Both halves carry the sequence number, so a reader that sees
// it in both has all four fields whether or not the 256-bit access is performed as one transaction.
__device__ __forceinline__ void lu_st_header(double2* p, double x0, double packed, double x1, double seq) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x0), "d"(packed), "d"(x1), "d"(seq) : "memory");
}
__device__ __forceinline__ void lu_ld_header(const double2* p, double& x0, double& packed, double& x1, double& seq) {
    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x0), "=d"(packed), "=d"(x1), "=d"(seq) : "l"(p) : "memory");
}

// 32-byte row word {v0, v1, v2, (double)seq}: three consecutive entries of a published candidate row
__device__ __forceinline__ void lu_st_row3(double2* p, double v0, double v1, double v2, double seq) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v0), "d"(v1), "d"(v2), "d"(seq) : "memory");
}
__device__ __forceinline__ void lu_ld_row3(const double2* p, double& v0, double& v1, double& v2, double& seq) {
    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v0), "=d"(v1), "=d"(v2), "=d"(seq) : "l"(p) : "memory");
}

// warp_best with a short path: when a single lane holds the largest high key word (the usual case: distinct
// |values|) it is the winner and its fields are fetched with independent shuffles -- one redux + one vote +
// shuffles instead of four dependent redux operations.
__device__ __forceinline__ void warp_best_fast(double& v, int& r, int& tag) {
    const unsigned FULL = 0xffffffffu;
    // keys are |x| >= 0, +inf, or the markers -1 / -2: hi32 + 2 / 1 / 0 is monotone in the key
    const unsigned hi = v >= 0.0 ? (unsigned)__double2hiint(v) + 2u : (v == -1.0 ? 1u : 0u);
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned m = __ballot_sync(FULL, hi == mhi);
    if (__popc(m) == 1) {
        const int src = __ffs(m) - 1;
        v = __shfl_sync(FULL, v, src); r = __shfl_sync(FULL, r, src); tag = __shfl_sync(FULL, tag, src);
    } else {
        warp_best(v, r, tag);
    }
}

// y + np * l with two roundings (unfused, like the reference's axpy), as ONE opaque operation: written as separate
// __dmul_rn / __dadd_rn the compiler hoists all the products of a window ahead of the adds and spills.
__device__ __forceinline__ double mul_add_unfused(double np, double l, double y) {
    double r;
    asm("{\n\t.reg .f64 t;\n\tmul.rn.f64 t, %1, %2;\n\tadd.rn.f64 %0, t, %3;\n\t}" : "=d"(r) : "d"(np), "d"(l), "d"(y));
    return r;
}

template <int W, int RPT>
struct Getf2Reg {
    static constexpr int NW3 = (W + 2) / 3;            // 32-byte row words per published row (entries 2.. of the window)
    static constexpr int SL = 2 * (1 + NW3);           // slot, in 16-byte units: header + row words (32 bytes each)
    static constexpr int ROWS = kRegThreads * RPT;

    struct Smem {
        // Pivot row of column c, entries c.., double-buffered by column parity (the next row is written while slow
        // warps still read this one), and kept TWICE: prow_a[i] = entry i, prow_b[i + 1] = entry i.  The bulk update
        // reads entries in pairs (c+3, c+4), (c+5, c+6), ...: whichever copy makes the first pair 16-byte aligned is
        // used, so one code path serves both parities of c.
        alignas(16) double prow_a[2][2 * W + 2];
        alignas(16) double prow_b[2][2 * W + 2];
        alignas(16) double zero_row[2 * W + 2];  // stands in for the pivot row of a skipped (all-zero) column
        double stage[2][W];                     // G == 1: the candidate row, written by its owner
        double ret_u[W][W];                     // U rows of the pivot rows this CTA owned (row c: entries c..)
        double lbuf[W][ROWS];                   // finished multipliers: lbuf[c][row] = L(row, c)
        double red_v[8]; int red_r[8], red_t[8];                        // per-warp local candidates
        double gred_v[8], gred_x0[8], gred_x1[8], gred_inv[8]; int gred_r[8], gred_c[8];   // per-warp partial reductions of the G headers
    };

    const Getf2RegParams& p;
    Smem& sm;
    // Register window: x[i][k] is the entry of row i in column (c + k) while column c is being eliminated; it is
    // shifted left by one per column inside the update itself (x[k] <- x[k+1] - prow[c+1+k] * l), so the column
    // loop is a real loop over 8 code blocks (window 64, 56, ...).  Measured alternatives, both slower: a fully
    // unrolled column loop (static register indices, no shift: instruction-fetch bound) and a window that stays
    // put for 8 columns with the column selected by a switch (more code, more barrier skew).
    double (&x)[RPT][W];
    int (&pos)[RPT];
    const int tid, lane, warp, G, cta, ncol;
    // winner of the local candidate search of the column about to be received (G == 1: the global winner)
    double lv; int lpos, ltag;
    long long t_prev = 0, t_acc[6] = {0, 0, 0, 0, 0, 0};

    __device__ __forceinline__ Getf2Reg(const Getf2RegParams& p_, Smem& sm_, double (&x_)[RPT][W], int (&pos_)[RPT])
        : p(p_), sm(sm_), x(x_), pos(pos_), tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5),
          G(gridDim.x), cta(blockIdx.x), ncol(min(p_.w, p_.m)), lv(-2.0), lpos(kNoRow), ltag(kNoRow) {}

    __device__ __forceinline__ double2* slot(int col, int c) const { return p.xch + ((size_t)(col & 1) * G + c) * SL; }

    // Local candidate of column cn (held in x[.][0]) over the live rows -> per-warp winners in shared memory.
    __device__ __forceinline__ void cand_local(int cn) {
        double bv = -2.0; int br = kNoRow, bt = kNoRow;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int pi = pos[i];
            const double key = pi >= 0 ? pivot_key(x[i][0], pi == cn) : -2.0;
            if (cand_better(key, pi, bv, br)) { bv = key; br = pi; bt = i * kRegThreads + tid; }
        }
        warp_best_fast(bv, br, bt);
        if (lane == 0) { sm.red_v[warp] = bv; sm.red_r[warp] = br; sm.red_t[warp] = bt; }
    }
    // After a block barrier: every warp reduces the 8 per-warp winners itself (no second barrier).
    __device__ __forceinline__ void cand_reduce() {
        lv = lane < 8 ? sm.red_v[lane] : -2.0; lpos = lane < 8 ? sm.red_r[lane] : kNoRow; ltag = lane < 8 ? sm.red_t[lane] : kNoRow;
        warp_best_fast(lv, lpos, ltag);
    }
    __device__ __forceinline__ bool is_owner() const { return lv != -2.0 && (ltag & (kRegThreads - 1)) == tid; }

    // Header of column cn: the owner of the local candidate sends the candidate's entries in columns cn and cn+1
    // (window entries 0 and 1, both final); a CTA without live rows says so (pos = kNoRow).
    __device__ __forceinline__ void publish_header(int cn) {
        const int iseq = p.seq0 + cn + 1;
        if (lv == -2.0) {
            if (tid == 0) lu_st_header(slot(cn, cta), 0.0, pack_seq_row(iseq, kNoRow), 0.0, (double)iseq);
            return;
        }
        if (!is_owner()) return;
        const int oi = ltag / kRegThreads;
#pragma unroll
        for (int i = 0; i < RPT; ++i)
            if (i == oi) lu_st_header(slot(cn, cta), x[i][0], pack_seq_row(iseq, lpos), W > 1 ? x[i][W > 1 ? 1 : 0] : 0.0, (double)iseq);
    }

    // The owner of the local candidate publishes the rest of its row.  G > 1: window entries 2.. (columns cn+2..; the
    // header carries entries 0 and 1) as 32-byte words of three entries + the sequence number.  G == 1: entries 0..
    // into the staging row.  Entries beyond column W-1 are padding; whole words beyond it are not sent.
    template <int WN>
    __device__ __forceinline__ void publish_row(int cn) {
        if (!is_owner()) return;
        const int oi = ltag / kRegThreads;
        const double seq = (double)(p.seq0 + cn + 1);
        double2* my = slot(cn, cta) + 2;
        double* st = sm.stage[cn & 1];
        const int nvalid = W - cn;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            if (i != oi) continue;
            if (G > 1) {
#pragma unroll
                for (int g = 0; 2 + 3 * g < WN; ++g) {
                    if (2 + 3 * g >= nvalid) break;
                    const int k = 2 + 3 * g;
                    lu_st_row3(my + 2 * g, x[i][k], k + 1 < WN ? x[i][k + 1 < WN ? k + 1 : k] : 0.0, k + 2 < WN ? x[i][k + 2 < WN ? k + 2 : k] : 0.0, seq);
                }
            } else {
#pragma unroll
                for (int k = 0; k < WN; ++k) {
                    if (k >= nvalid) break;
                    st[cn + k] = x[i][k];
                }
            }
        }
    }

    // Receives column c: the global winner (position gpos, CTA gcta), its entries x0 (column c: the pivot) and x1
    // (column c+1), and inv = 1/x0 (IEEE-rounded, as gauss_step computes it; 1 when the column is skipped).  G > 1: from
    // the G headers, one exchange; every polling thread also takes the reciprocal of the header it read, so the division
    // overlaps the reductions instead of following them.
    __device__ __forceinline__ void receive(int c, int& gpos, int& gcta, double& x0, double& x1, double& inv) {
        if (G > 1) {
            const unsigned FULL = 0xffffffffu;
            const int iseq = p.seq0 + c + 1;
            const double dseq = (double)iseq;
            double gv = -2.0, h0 = 0.0, h1 = 0.0, hi = 1.0; int gr = kNoRow, gc = kNoRow;
            if (tid < G) {
                const double2* h = slot(c, tid);
                double packed, s2;
                do { lu_ld_header(h, h0, packed, h1, s2); } while ((int)(__double_as_longlong(packed) >> 32) != iseq || s2 != dseq);
                gr = (int)(__double_as_longlong(packed) & 0xffffffffLL); gc = tid;
                gv = gr == kNoRow ? -2.0 : pivot_key(h0, gr == c);
                hi = h0 != 0.0 ? __drcp_rn(h0) : 1.0;
            }
            const int nw = (G + 31) >> 5;
            if (warp < nw) {
                int src = lane;
                warp_best_fast(gv, gr, src);                  // src = lane that read the winning header
                src &= 31;
                h0 = __shfl_sync(FULL, h0, src); h1 = __shfl_sync(FULL, h1, src); hi = __shfl_sync(FULL, hi, src); gc = __shfl_sync(FULL, gc, src);
                if (lane == 0) { sm.gred_v[warp] = gv; sm.gred_r[warp] = gr; sm.gred_c[warp] = gc; sm.gred_x0[warp] = h0; sm.gred_x1[warp] = h1; sm.gred_inv[warp] = hi; }
            }
            __syncthreads();
            gv = lane < nw ? sm.gred_v[lane] : -2.0; gr = lane < nw ? sm.gred_r[lane] : kNoRow;
            int src = lane;
            warp_best_fast(gv, gr, src);                      // every warp reduces the <= 5 partial winners itself
            src &= 7;
            gpos = gr; gcta = sm.gred_c[src]; x0 = sm.gred_x0[src]; x1 = sm.gred_x1[src]; inv = sm.gred_inv[src];
        } else {
            __syncthreads();                                  // the owner's publish_row wrote the staging row
            gpos = lpos; gcta = 0; x0 = sm.stage[c & 1][c]; x1 = c + 1 < W ? sm.stage[c & 1][c + 1 < W ? c + 1 : c] : 0.0;
            inv = x0 != 0.0 ? __drcp_rn(x0) : 1.0;
        }
    }

    // One column; the window holds WIN entries (columns c .. c+WIN-1; those beyond W-1 are padding).
    template <int WIN>
    __device__ __forceinline__ void step(int c) {
        int gpos, gcta;
        double x0, x1, inv;
        RPROF(0);
        receive(c, gpos, gcta, x0, x1, inv);
        RPROF(1);
        // Rest of the winner's row (window entries 2.. = columns c+2..): requested now, stored to the two prow copies just
        // before the next barrier.  G > 1: thread g fetches the 32-byte word g; G == 1: thread t copies staging entry c + t.
        constexpr int NG = WIN > 2 ? (WIN - 2 + 2) / 3 : 0;
        const int nvalid = W - c;                            // real (non-padding) window entries
        const bool rload = G > 1 && tid < NG && 2 + 3 * tid < nvalid;
        const double2* rp = slot(c, gcta) + 2 + 2 * tid;
        const double dseq = (double)(p.seq0 + c + 1);
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, rs = 0.0;
        if (rload) lu_ld_row3(rp, r0, r1, r2, rs);
        const bool elim = x0 != 0.0;                         // lu.rs:107-110: an all-zero column is skipped (then gpos == c)
        if (cta == 0 && tid == 0) p.ipiv[p.j0 + c] = p.j0 + (elim ? gpos : c);
        // The winner retires into position c: its U entries are the pivot row everybody holds (kept by the owning
        // CTA in ret_u), its L entries are already in lbuf.  The row that sat at position c moves to the winner's
        // old position.  Retired (and padding) rows keep being "updated" below -- their window is dead, so the
        // update needs no predicate.  A skipped column runs the same code with inv = 1 and a zero pivot row
        // (x + 0 * l = x; only the sign of a zero may differ from an untouched entry).
        double l[RPT];
        const double nx1 = elim ? -x1 : 0.0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int pi = pos[i];
            pos[i] = pi == gpos ? -2 - c : (pi == c ? gpos : pi);
            l[i] = __dmul_rn(x[i][0], inv);
            sm.lbuf[c][i * kRegThreads + tid] = l[i];
            // column c+1 first (unfused mul/add like the reference): its candidates leave before the bulk update
            if constexpr (WIN > 1) x[i][0] = __dadd_rn(__dmul_rn(nx1, l[i]), x[i][1]);
        }
        const bool more = WIN > 1 && c + 1 < ncol;
        RPROF(2);
        if (more) cand_local(c + 1);
        double* pa = sm.prow_a[c & 1];
        double* pb = sm.prow_b[c & 1];
        if (G > 1) {
            if (rload) {
                while (rs != dseq) lu_ld_row3(rp, r0, r1, r2, rs);
                const int i0 = c + 2 + 3 * tid;
                pa[i0] = r0; pb[i0 + 1] = r0;
                pa[i0 + 1] = r1; pb[i0 + 2] = r1;             // entries beyond W-1 land in the padding of the 2W + 2 arrays
                pa[i0 + 2] = r2; pb[i0 + 3] = r2;
            }
            if (tid == 0) { pa[c] = x0; pb[c + 1] = x0; pa[c + 1] = x1; pb[c + 2] = x1; }
        } else if (tid < nvalid) {
            const double v = sm.stage[c & 1][c + tid];
            pa[c + tid] = v; pb[c + tid + 1] = v;
        }
        __syncthreads();
        if (more) cand_reduce();
        const double* prow = elim ? pa + c : sm.zero_row;    // prow[k] = pivot row entry of column c + k
        if constexpr (WIN > 2) {
            const double p2 = -prow[2];
#pragma unroll
            for (int i = 0; i < RPT; ++i) x[i][1] = __dadd_rn(__dmul_rn(p2, l[i]), x[i][2]);
        }
        if (more && G > 1) publish_header(c + 1);
        RPROF(3);
        if (gcta == cta && tid >= c && tid < W) sm.ret_u[c][tid] = pa[tid];
        if constexpr (WIN > 3) {
            // x[k] <- x[k+1] - prow[k+1] * l, k = 2 .. WIN-2, the pivot row read in aligned pairs (entries c+3, c+4), ...
            // from the copy whose alignment fits: prow_a when c + 3 is even, prow_b (shifted by one) when it is odd
            const double2* pp = reinterpret_cast<const double2*>(!elim ? sm.zero_row : ((c & 1) ? pa + c + 3 : pb + c + 4));
#pragma unroll
            for (int k = 2; k + 1 < WIN; k += 2) {
                const double2 pv = pp[(k - 2) >> 1];
#pragma unroll
                for (int i = 0; i < RPT; ++i) x[i][k] = mul_add_unfused(-pv.x, l[i], x[i][k + 1]);
                if (k + 2 < WIN) {
#pragma unroll
                    for (int i = 0; i < RPT; ++i) x[i][k + 1] = mul_add_unfused(-pv.y, l[i], x[i][k + 2]);
                }
            }
        }
        RPROF(4);
        if (more) publish_row<(WIN > 1 ? WIN - 1 : 1)>(c + 1);
        RPROF(5);
    }

    // Columns [W - WIN, W - WIN + kRegBlock) with a window of WIN entries, then the next block.
    template <int WIN>
    __device__ __forceinline__ void run() {
        const int c0 = W - WIN;
        if (c0 >= ncol) return;
        const int c1 = min(ncol, c0 + kRegBlock);
        for (int c = c0; c < c1; ++c) step<WIN>(c);
        if constexpr (WIN > kRegBlock) run<WIN - kRegBlock>();
    }
};

template <int W, int RPT>
__global__ void __launch_bounds__(kRegThreads, 1) getf2_reg_kernel(const Getf2RegParams p) {
    using K = Getf2Reg<W, RPT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename K::Smem& sm = *reinterpret_cast<typename K::Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int r_begin = blockIdx.x * p.rows_cta;
    const int nrows = max(0, min(p.rows_cta, p.m - r_begin));
    for (int k = tid; k < 2 * W + 2; k += kRegThreads) {
        sm.prow_a[0][k] = 0.0; sm.prow_a[1][k] = 0.0; sm.prow_b[0][k] = 0.0; sm.prow_b[1][k] = 0.0; sm.zero_row[k] = 0.0;
    }
    __syncthreads();
    double x[RPT][W];
    int pos[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int lr = i * kRegThreads + tid;
        const bool valid = lr < nrows;
        pos[i] = valid ? r_begin + lr : kInvalidRow;
        const double* src = p.a + (long long)(r_begin + lr);
#pragma unroll
        for (int cc = 0; cc < W; ++cc) x[i][cc] = (valid && cc < p.w) ? src[(long long)cc * p.lda] : 0.0;
    }
    K k(p, sm, x, pos);
#ifdef NAB_GETF2_PROF
    k.t_prev = clock64();
#endif
    k.cand_local(0);
    __syncthreads();
    k.cand_reduce();
    if (gridDim.x > 1) k.publish_header(0);
    k.template publish_row<W>(0);
    k.template run<W>();
#ifdef NAB_GETF2_PROF
    if (tid == 0 && blockIdx.x == 0)
        for (int i = 0; i < 6; ++i) g_getf2_reg_prof[i] += k.t_acc[i];
#endif
    __syncthreads();
    // every row goes to its final position.  A live row (never a pivot; only when the panel has more rows than
    // columns) has all its entries in lbuf; a row retired at column c has its L entries (columns < c) in lbuf and
    // its U entries in ret_u[c].  When the panel is wider than tall (ncol = m < w) every row is retired.
    const int w = p.w;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int pi = pos[i];
        if (pi == kInvalidRow) continue;
        const int cret = pi >= 0 ? w : -2 - pi;
        const int pf = pi >= 0 ? pi : cret;                  // final position of this row
        if (p.list != nullptr && pf != r_begin + i * kRegThreads + tid) {
            const int idx = atomicAdd(p.list, 1);
            p.list[1 + idx] = p.j0 + pf;
            p.list[1 + kRegListMax + idx] = p.j0 + r_begin + i * kRegThreads + tid;
        }
        double* dst = p.a + (long long)pf;
        const double* ur = sm.ret_u[cret < W ? cret : 0];
        const int row = i * kRegThreads + tid;
        for (int cc = 0; cc < w; ++cc) dst[(long long)cc * p.lda] = cc < cret ? sm.lbuf[cc][row] : ur[cc];
    }
}

template <int W, int RPT>
static int launch_reg(cudaStream_t st, const Getf2RegParams& p, int G) {
    using K = Getf2Reg<W, RPT>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(getf2_reg_kernel<W, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(typename K::Smem));
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(getf2_reg)", __FILE__, __LINE__);
    void* args[] = {(void*)&p};
    NAB_CUDA(cudaLaunchCooperativeKernel((void*)getf2_reg_kernel<W, RPT>, dim3((unsigned)G), dim3(kRegThreads), args,
                                         sizeof(typename K::Smem), st));
    count_launch();
    return NA_OK;
}

// Rows per CTA of the register-resident leaf for a panel of width w.
int getf2_reg_rows_per_cta(size_t w) {
    static const int force_w = [] { const char* e = getenv("NAB_GETF2_W"); return e ? atoi(e) : 0; }();
    if (force_w == 64) return kRegThreads;
    return w <= 16 ? 4 * kRegThreads : (w <= 32 ? 2 * kRegThreads : kRegThreads);
}

// CTAs the register-resident leaf needs for an m x w panel (0: the panel does not fit this kernel).
int getf2_reg_grid(size_t m, size_t w) {
    if (w == 0 || w > 64 || m == 0) return 0;
    const size_t G = ceil_div(m, (size_t)getf2_reg_rows_per_cta(w));
    return G <= (size_t)ctx().sm_count ? (int)G : 0;
}

// Factors the m x w panel at A[j0.., j0..j0+w) (w <= 64).  ws / seq_state: the workspace and sequence counter of
// getf2_panel (the two kernels may alternate on the same workspace: sequence numbers only grow).
int getf2_panel_reg(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, size_t j0, int* ipiv, void* ws, int* seq_state,
                    int cta_limit, int* list) {
    if (m == 0 || w == 0) return NA_OK;
    const int G0 = getf2_reg_grid(m, w);
    if (G0 == 0) { set_error("getf2_reg: a %zu x %zu panel does not fit the register-resident leaf", m, w); return NA_EINVAL; }
    if (cta_limit > 0 && G0 > cta_limit) { set_error("getf2_reg: needs %d CTAs, %d allowed", G0, cta_limit); return NA_EINVAL; }
    // spread the rows evenly over the CTAs (multiples of 32 rows)
    const size_t rows_cta = round_up(ceil_div(m, (size_t)G0), 32);
    const int G = (int)ceil_div(m, rows_cta);
    Getf2RegParams p;
    p.a = a_panel; p.lda = (long long)lda; p.m = (int)m; p.w = (int)w; p.rows_cta = (int)rows_cta; p.j0 = (int)j0; p.ipiv = ipiv;
    p.xch = static_cast<double2*>(ws);
    p.seq0 = *seq_state;
    p.list = list;
    *seq_state += (int)w + 2 + ((w & 1) ? 1 : 0);
    static const int force_w = [] { const char* e = getenv("NAB_GETF2_W"); return e ? atoi(e) : 0; }();   // timing experiments
    if (force_w == 64 && (size_t)G * kRegThreads >= m) return launch_reg<64, 1>(st, p, G);
    if (w <= 16) return launch_reg<16, 4>(st, p, G);
    if (w <= 32) return launch_reg<32, 2>(st, p, G);
    return launch_reg<64, 1>(st, p, G);
}

}  // namespace nab

#ifdef NAB_GETF2_PROF
extern "C" __attribute__((visibility("default"))) int na_debug_getf2_reg_prof(long long* out, int reset) {
    cudaMemcpyFromSymbol(out, nab::g_getf2_reg_prof, sizeof(long long) * 16);
    if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(nab::g_getf2_reg_prof, z, sizeof(z)); }
    return 0;
}
#endif
