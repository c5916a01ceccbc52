// factor_lu.cu -- recursive blocked LU with partial pivoting and LU solve, on the DGEMM engine.
//
// Reference: LU::new (/root/reference/src/linalg/lu.rs:93-122), gauss_step(_swap) (:337-389),
// LU::solve_mut (:242-260), PermutationSequence (src/linalg/permutation_sequence.rs).
// The reference is a right-looking rank-1 loop; here the columns are split recursively
// (left half -> row swaps + TRSM + GEMM on the right half -> right half -> row swaps on the left
// half), so all O(n^3) work is GEMM/TRSM on the DMMA engine and the leaves are the cooperative
// GETF2 panel kernel.  Whole rows are swapped, so the packed result has the reference's layout
// (identical to LAPACK getrf).  Everything stays on the device: no host pivoting.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"

namespace nab {

struct LuCtx {
    cudaStream_t s;
    double* a; size_t lda; size_t M;
    size_t W;                 // leaf panel width
    int* ipiv;                // device, absolute pivot rows
    int* iota;                // device, 0..mn-1
    void* ws_getf2; void* ws_perm;
    int* seq_state;           // host counter of exchange sequence numbers used in ws_getf2
    int getf2_limit = 0;      // max CTAs of a GETF2 leaf (look-ahead: the SMs kept free of the bulk GEMM); 0 = no limit
    Timeline* tl = nullptr;   // optional per-node timeline of the panel recursion
    int* lists = nullptr;     // device: one moved-row list (kRegListInts ints, count zeroed) per leaf of this factorization
    int* leaf_seq = nullptr;  // host counter: lists handed out so far
    size_t max_leaves = 0;
};

// NAB_GETF2=smem selects the shared-memory leaf of panel_lu.cu everywhere (A/B timing); default: the
// register-resident leaf wherever the panel fits it.
static bool lu_use_reg() {
    static bool v = [] { const char* e = getenv("NAB_GETF2"); return !(e && strcmp(e, "smem") == 0); }();
    return v;
}

// One GETF2 leaf.  *list (when the register-resident leaf ran and a list slot was left) receives the leaf's
// moved-row list for rowperm_apply_lists; nullptr means "build the permutation from ipiv".
static int lu_leaf(const LuCtx& c, double* ajj, size_t m, size_t w, size_t j0, int** list = nullptr) {
    if (list) *list = nullptr;
    const int greg = lu_use_reg() ? getf2_reg_grid(m, w) : 0;
    if (greg > 0 && (c.getf2_limit == 0 || greg <= c.getf2_limit)) {
        int* l = nullptr;
        if (list && c.lists && (size_t)*c.leaf_seq < c.max_leaves) { l = c.lists + (size_t)(*c.leaf_seq)++ * kRegListInts; *list = l; }
        return getf2_panel_reg(c.s, ajj, c.lda, m, w, j0, c.ipiv, c.ws_getf2, c.seq_state, c.getf2_limit, l);
    }
    return getf2_panel(c.s, ajj, c.lda, m, w, j0, c.ipiv, c.ws_getf2, c.seq_state, c.getf2_limit);
}

static int lu_apply_swaps(const LuCtx& c, size_t k0, size_t K, double* cols, size_t ncols) {
    if (K == 0 || ncols == 0) return NA_OK;
    NAB_TRY(rowperm_build(c.s, c.iota + k0, c.ipiv + k0, K, 1, c.M, c.ws_perm));
    return rowperm_apply(c.s, cols, c.lda, ncols, std::min(2 * K, c.M), c.ws_perm, c.M);
}

// Panels of up to lu_flat() columns: flat right-looking loop over W-column leaves --
//   GETF2 leaf -> its row moves on the columns right of it (inside the panel) -> U12 = L11^-1 A12 (direct small TRSM)
//   -> A22 -= A21 U12 (one rank-W GEMM) -> next leaf; the leaves' row moves on the columns LEFT of them come last.
// 4 launches per leaf instead of the ~7 of the recursive split (no perm build: the register-resident leaf emits its
// moved-row list; one small TRSM per leaf instead of TRSM + GEMM + TRSM chains at the upper nodes).
static size_t lu_flat() {
    static size_t v = [] { const char* e = getenv("NAB_LU_FLAT"); return e ? (size_t)atoi(e) : (size_t)512; }();
    return v;
}

static bool lu_rank_update_enabled() {
    static bool v = [] { const char* e = getenv("NAB_LU_RANKK"); return e ? atoi(e) != 0 : true; }();
    return v;
}

static int lu_swaps_from_leaf(const LuCtx& c, const int* list, size_t k0, size_t K, double* cols, size_t ncols) {
    if (ncols == 0 || K == 0) return NA_OK;
    if (list) return rowperm_apply_lists(c.s, cols, c.lda, ncols, std::min<size_t>(2 * K, kRegListMax), list, list + 1, list + 1 + kRegListMax);
    return lu_apply_swaps(c, k0, K, cols, ncols);
}

static int lu_panel_flat(const LuCtx& c, size_t j0, size_t nc) {
    std::vector<int*> lists;
    std::vector<size_t> starts;
    for (size_t k0 = 0; k0 < nc; k0 += c.W) {
        const size_t wk = std::min(c.W, nc - k0), jk = j0 + k0, mk = c.M - jk;
        double* akk = c.a + jk + jk * c.lda;
        int* list = nullptr;
        cudaEvent_t t0 = c.tl ? c.tl->mark(c.s) : nullptr;
        NAB_TRY(lu_leaf(c, akk, mk, wk, jk, &list));
        cudaEvent_t t1 = c.tl ? c.tl->mark(c.s) : nullptr;
        lists.push_back(list); starts.push_back(k0);
        const size_t n2 = nc - k0 - wk, kk = std::min(wk, mk);
        if (n2 > 0) {
            double* a12 = akk + wk * c.lda;
            NAB_TRY(lu_swaps_from_leaf(c, list, jk, kk, a12 - jk, n2));                  // whole rows: columns start at row 0
            cudaEvent_t t2 = c.tl ? c.tl->mark(c.s) : nullptr;
            NAB_TRY(trsm_unit_lower_small(c.s, kk, akk, c.lda, a12, c.lda, n2));
            cudaEvent_t t3 = c.tl ? c.tl->mark(c.s) : nullptr;
            if (mk > wk) {
                // rank-wk update of the rest of the panel: streaming DMMA kernel (K <= 64 is a bandwidth problem), GEMM engine otherwise
                int ru = lu_rank_update_enabled() ? rank_update_small_k(c.s, mk - wk, wk, n2, akk + wk, c.lda, a12, c.lda, a12 + wk, c.lda, c.getf2_limit) : 1;
                if (ru < 0) return ru;
                if (ru == 1)
                    NAB_TRY(dgemm_device(c.s, false, mk - wk, wk, n2, -1.0, akk + wk, 1, (ptrdiff_t)c.lda, a12, 1, (ptrdiff_t)c.lda, 1.0,
                                         a12 + wk, 1, (ptrdiff_t)c.lda));
            }
            if (c.tl) {
                cudaEvent_t t4 = c.tl->mark(c.s);
                c.tl->add("f.swap", jk, t1, t2); c.tl->add("f.trsm", jk, t2, t3); c.tl->add("f.gemm", jk, t3, t4);
            }
        }
        if (c.tl) c.tl->add("getf2", jk, t0, t1);
    }
    cudaEvent_t t5 = c.tl ? c.tl->mark(c.s) : nullptr;
    for (size_t i = 1; i < lists.size(); ++i) {
        const size_t k0 = starts[i], jk = j0 + k0;
        NAB_TRY(lu_swaps_from_leaf(c, lists[i], jk, std::min(std::min(c.W, nc - k0), c.M - jk), c.a + j0 * c.lda, k0));
    }
    if (c.tl) c.tl->add("f.swapL", j0, t5, c.tl->mark(c.s));
    return NA_OK;
}

static int lu_rec(const LuCtx& c, size_t j0, size_t nc) {
    if (nc == 0) return NA_OK;
    double* ajj = c.a + j0 + j0 * c.lda;
    if (nc > c.W && nc <= lu_flat() && c.W <= 64) return lu_panel_flat(c, j0, nc);
    if (nc <= c.W) {
        cudaEvent_t t = c.tl ? c.tl->mark(c.s) : nullptr;
        NAB_TRY(lu_leaf(c, ajj, c.M - j0, nc, j0));
        if (c.tl) c.tl->add("getf2", j0, t, c.tl->mark(c.s));
        return NA_OK;
    }
    size_t n1 = round_up(nc / 2, c.W);
    if (n1 >= nc) n1 = nc - c.W;
    const size_t n2 = nc - n1;
    NAB_TRY(lu_rec(c, j0, n1));
    double* a12 = ajj + n1 * c.lda;
    cudaEvent_t tn0 = c.tl ? c.tl->mark(c.s) : nullptr;
    NAB_TRY(lu_apply_swaps(c, j0, n1, a12 - j0, n2));                         // whole rows: columns start at row 0
    cudaEvent_t tn1 = c.tl ? c.tl->mark(c.s) : nullptr;
    // U12 = L11^-1 * A12 (unit lower): direct substitution kernels for the narrow nodes of the panel recursion
    // (n1 <= 128; 256 = two of them around one GEMM), blocked inverse-based TRSM above that
    if (n1 <= 128) {
        NAB_TRY(trsm_unit_lower_small(c.s, n1, ajj, c.lda, a12, c.lda, n2));
    } else if (n1 <= 256) {
        const size_t h = 128, r = n1 - h;
        NAB_TRY(trsm_unit_lower_small(c.s, h, ajj, c.lda, a12, c.lda, n2));
        NAB_TRY(dgemm_device(c.s, false, r, h, n2, -1.0, ajj + h, 1, (ptrdiff_t)c.lda, a12, 1, (ptrdiff_t)c.lda, 1.0, a12 + h, 1, (ptrdiff_t)c.lda));
        NAB_TRY(trsm_unit_lower_small(c.s, r, ajj + h + h * c.lda, c.lda, a12 + h, c.lda, n2));
    } else {
        NAB_TRY(trsm_left(c.s, true, true, n1, ajj, 1, (ptrdiff_t)c.lda, nullptr, nullptr, a12, 1, (ptrdiff_t)c.lda, n2));
    }
    // A22 -= A21 * U12
    cudaEvent_t tn2 = c.tl ? c.tl->mark(c.s) : nullptr;
    const size_t m2 = c.M - j0 - n1;
    if (m2 > 0)
        NAB_TRY(dgemm_device(c.s, false, m2, n1, n2, -1.0, ajj + n1, 1, (ptrdiff_t)c.lda, a12, 1, (ptrdiff_t)c.lda, 1.0,
                             a12 + n1, 1, (ptrdiff_t)c.lda));
    if (c.tl) {
        cudaEvent_t tn3 = c.tl->mark(c.s);
        c.tl->add("n.swap", j0 + n1, tn0, tn1); c.tl->add("n.trsm", j0 + n1, tn1, tn2); c.tl->add("n.gemm", j0 + n1, tn2, tn3);
    }
    NAB_TRY(lu_rec(c, j0 + n1, n2));
    cudaEvent_t tn4 = c.tl ? c.tl->mark(c.s) : nullptr;
    NAB_TRY(lu_apply_swaps(c, j0 + n1, std::min(n2, c.M - j0 - n1), c.a + j0 * c.lda, n1));
    if (c.tl) c.tl->add("n.swapL", j0 + n1, tn4, c.tl->mark(c.s));
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// Large square-ish matrices: right-looking outer blocks of LU_NB columns with one-step look-ahead.
//   panel(j)  : lu_rec on columns [j, j+nb) (leaves of 64 columns -> the cooperative GETF2 needs
//               <= 43 SMs), on the caller's stream, GEMMs limited to `rp` CTAs;
//   la(j)     : row swaps + TRSM + GEMM of panel j on the NEXT panel's columns, whole GPU;
//   bulk(j)   : the same on all remaining columns (and the row swaps on the columns left of the
//               panel), on a second stream with the remaining SMs, concurrent with panel(j + nb).
// ------------------------------------------------------------------------------------------------
static size_t lu_nb() {
    static size_t v = [] { const char* e = getenv("NAB_LU_NB"); size_t x = e ? (size_t)atoi(e) : 512; return x >= 128 ? x / 64 * 64 : 512; }();
    return v;
}
#define LU_NB (lu_nb())

static int lu_right_update(const LuCtx& c, cudaStream_t st, size_t j, size_t jb, size_t x0, size_t nx, const void* ws_outer,
                           const double* inv_l11) {
    if (nx == 0) return NA_OK;
    double* ax = c.a + x0 * c.lda;                      // column x0, row 0
    NAB_TRY(rowperm_apply(st, ax, c.lda, nx, std::min(2 * jb, c.M), ws_outer, c.M));
    NAB_TRY(trsm_left(st, true, true, jb, c.a + j + j * c.lda, 1, (ptrdiff_t)c.lda, nullptr, inv_l11, ax + j, 1, (ptrdiff_t)c.lda, nx));
    const size_t m2 = c.M - j - jb;
    if (m2 > 0)
        NAB_TRY(dgemm_device(st, false, m2, jb, nx, -1.0, c.a + (j + jb) + j * c.lda, 1, (ptrdiff_t)c.lda, ax + j, 1, (ptrdiff_t)c.lda,
                             1.0, ax + j + jb, 1, (ptrdiff_t)c.lda));
    return NA_OK;
}

// most SMs the panel chain may take once the whole bulk update fits beside it (the in-panel updates scale with it)
static int lu_rp_cap() {
    static int v = [] { const char* e = getenv("NAB_LU_RPCAP"); return e ? atoi(e) : 112; }();
    return v;
}

static bool lu_split() {
    static bool v = [] { const char* e = getenv("NAB_LU_SPLIT"); return e ? atoi(e) != 0 : true; }();
    return v;
}

static int lu_lookahead(LuCtx c, size_t N, size_t mn) {
    // The panel chain (la + panel) runs on an internal HIGH-priority stream and the bulk update on a normal one:
    // when both have CTAs pending (right after la, when bulk(j)'s row swaps / TRSM flood the SMs) the cooperative
    // GETF2 launch of the next panel's first leaf is placed first instead of waiting ~100 us behind them.
    StreamGuard sp_g, su_g;
    EventGuard ev_p, ev_u, ev_d, ev_in;
    NAB_TRY(sp_g.create(true)); NAB_TRY(su_g.create(false));
    NAB_TRY(ev_p.create()); NAB_TRY(ev_u.create()); NAB_TRY(ev_d.create()); NAB_TRY(ev_in.create());
    const cudaStream_t caller = c.s, sp = sp_g.s, su = su_g.s;
    NAB_CUDA(cudaEventRecord(ev_in, caller));
    NAB_CUDA(cudaStreamWaitEvent(sp, ev_in, 0));
    c.s = sp;
    Scratch wso[2], invb[2];      // per-step permutation lists and inverses of L11's diagonal blocks, double-buffered
    int st = wso[0].alloc(rowperm_workspace_bytes(c.M), sp);
    if (st == NA_OK) st = wso[1].alloc(rowperm_workspace_bytes(c.M), sp);
    const size_t inv_bytes = ceil_div(LU_NB, (size_t)kInvBlock) * kInvBlock * kInvBlock * sizeof(double);
    if (st == NA_OK) st = invb[0].alloc(inv_bytes, sp);
    if (st == NA_OK) st = invb[1].alloc(inv_bytes, sp);
    const int sms = ctx().sm_count;
    // CTAs (= SMs kept free of the bulk GEMM) the GETF2 leaf needs for a panel of m rows: the register-resident leaf
    // holds 256 rows of 64 columns per CTA, the shared-memory leaf 384
    const bool reg_leaf = lu_use_reg() && getf2_reg_grid(c.M, c.W) > 0;
    auto leaf_ctas = [&](size_t m) { return reg_leaf ? getf2_reg_grid(m, c.W) : (int)ceil_div(m, (size_t)384) + 2; };
    static const double tp0 = [] { const char* e = getenv("NAB_LU_TP0"); return e ? atof(e) : 0.0; }();
    static const double tp1 = [] { const char* e = getenv("NAB_LU_TP1"); return e ? atof(e) : 0.0; }();
    bool bulk_pending = false;
    int par = 0;
    Timeline tr("NAB_LU_TRACE", "lu_trace");
    tr.start(sp);
    if (tr.on && getenv("NAB_LU_TRACE_NODES")) c.tl = &tr;
    cudaEvent_t t_first = tr.mark(sp);
    if (st == NA_OK) st = lu_rec(c, 0, std::min(LU_NB, mn));
    tr.add("panel", 0, t_first, tr.mark(sp));
    for (size_t j = 0; st == NA_OK; j += LU_NB) {
        const size_t jb = std::min(LU_NB, mn - j), jn = j + jb;
        const size_t jbn = jn < mn ? std::min(LU_NB, mn - jn) : 0;
        st = rowperm_build(sp, c.iota + j, c.ipiv + j, jb, 1, c.M, wso[par].p);
        if (st == NA_OK) st = trtri_blocks(sp, c.a + j + j * c.lda, 1, (ptrdiff_t)c.lda, jb, true, true, nullptr, invb[par].as<double>());
        if (st != NA_OK) break;
        // la(j): the next panel's columns, whole GPU (needs bulk(j - nb) finished on them).  It is on the critical
        // chain, so bulk(j) is released only after it: side by side with part A it ran on the rp free SMs only.
        if (bulk_pending) cudaStreamWaitEvent(sp, ev_u, 0);
        cudaEvent_t t_la = tr.mark(sp);
        st = lu_right_update(c, sp, j, jb, jn, jbn, wso[par].p, invb[par].as<double>());
        if (st != NA_OK) break;
        tr.add("la", j, t_la, tr.mark(sp));
        cudaEventRecord(ev_p, sp);
        // bulk(j) on the second stream: left swaps + everything right of the next panel, in two parts.
        // Part A (the leftmost `wa` columns) runs on sms - rp CTAs next to panel(j + nb); part B (the rest)
        // starts when that panel is done and takes the whole GPU: the rp SMs reserved for the latency-bound
        // panel chain (~5.3 us per column whatever its height) idle only while the panel actually runs.
        const size_t x0 = jn + jbn, nx = N - x0, m2 = c.M - jn;
        // panel duration model (per column: leaf + in-panel updates), measured per 512 columns: shared-memory leaf
        // 2.3 .. 4.0 ms, register-resident leaf 1.06 ms (m = 512) .. 1.8 ms (m >= 7000, where the leaf is exchange-bound)
        const double t_panel = (double)jbn * (reg_leaf ? (tp0 > 0 ? tp0 : 2.0e-6) + (tp1 > 0 ? tp1 : 1.2e-6) * std::min(1.0, (double)m2 / 7000.0)
                                                       : 4.5e-6 + 2.5e-6 * (double)m2 / 16384.0);
        // SMs for the panel chain: what the GETF2 leaf needs at this height; more once the whole bulk update fits
        // beside the panel anyway.
        int rp = leaf_ctas(m2);
        {
            const double bulk_flops = 2.0 * (double)m2 * (double)jb * (double)nx;
            for (int r = rp; r <= std::min(lu_rp_cap(), sms - 16); r += 4)
                if (bulk_flops / ((sms - r) * kSmFlops) <= t_panel) rp = r;
        }
        size_t wa = nx;
        if (jbn && nx && m2 && lu_split()) {
            const double target = t_panel * (sms - rp) * kSmFlops;
            wa = round_up((size_t)(target / (2.0 * (double)m2 * (double)jb)) + 1, 128);
            if (wa + 256 >= nx) wa = nx;
        }
        cudaStreamWaitEvent(su, ev_p, 0);
        cudaEvent_t t_a = tr.mark(su);
        set_gemm_sm_limit(jbn ? sms - rp : 0);
        st = rowperm_apply(su, c.a, c.lda, j, std::min(2 * jb, c.M), wso[par].p, c.M);
        if (st == NA_OK) st = lu_right_update(c, su, j, jb, x0, wa, wso[par].p, invb[par].as<double>());
        set_gemm_sm_limit(0);
        if (st != NA_OK) break;
        tr.add("bulkA", j, t_a, tr.mark(su));
        if (jbn == 0) { cudaEventRecord(ev_u, su); bulk_pending = true; break; }
        // panel(j + nb) on the caller's stream, concurrently with part A
        cudaEvent_t t_p = tr.mark(sp);
        set_gemm_sm_limit(rp);
        c.getf2_limit = rp;
        st = lu_rec(c, jn, jbn);
        c.getf2_limit = 0;
        set_gemm_sm_limit(0);
        if (st != NA_OK) break;
        tr.add("panel", jn, t_p, tr.mark(sp));
        if (wa < nx) {
            cudaEventRecord(ev_d, sp);
            cudaStreamWaitEvent(su, ev_d, 0);
            cudaEvent_t t_b = tr.mark(su);
            st = lu_right_update(c, su, j, jb, x0 + wa, nx - wa, wso[par].p, invb[par].as<double>());
            if (st != NA_OK) break;
            tr.add("bulkB", j, t_b, tr.mark(su));
        }
        cudaEventRecord(ev_u, su);
        bulk_pending = true;
        par ^= 1;
    }
    if (bulk_pending) cudaStreamWaitEvent(sp, ev_u, 0);
    cudaEventRecord(ev_in, sp);
    cudaStreamWaitEvent(caller, ev_in, 0);             // the caller's stream continues after the whole factorization
    cudaStreamSynchronize(su);
    tr.dump();
    return st;
}

static size_t lu_leaf_width(size_t M) {
    if (lu_use_reg()) {                          // widest register-resident leaf whose rows fit the SMs
        for (size_t w : {64, 32, 16})
            if (getf2_reg_grid(M, w) > 0) return w;
    }
    if (M <= 28000) return 128;
    if (M <= 56000) return 64;
    if (M <= 113000) return 32;
    return 16;
}

static long g_lu_lookahead = 1;      // na_set_tuning("lu_lookahead", 0) forces the plain recursive / flat path
void lu_set_lookahead(long v) { g_lu_lookahead = v; }

// Everything but the pivot read-back: enqueues the factorization on `s`; ipiv_dev (device, min(M, N) ints) receives
// the 0-based pivot row of every column (LAPACK ipiv minus one; ipiv[i] == i where no swap happened).
int lu_device_async(cudaStream_t s, size_t M, size_t N, double* a, size_t lda, int* ipiv_dev) {
    const size_t mn = std::min(M, N);
    if (mn == 0) return NA_OK;
    if (lda < M) { set_error("lu: lda < m"); return NA_EINVAL; }
    if (M > 0x7fffff00ull || N > 0x7fffff00ull) { set_error("lu: dimension exceeds 2^31"); return NA_EINVAL; }
    Scratch iota, wsg, wsp;
    NAB_TRY(iota.alloc(mn * sizeof(int), s));
    NAB_TRY(wsg.alloc(getf2_workspace_bytes(), s));
    NAB_CUDA(cudaMemsetAsync(wsg.p, 0, getf2_workspace_bytes(), s));
    int seq_state = 0;
    NAB_TRY(wsp.alloc(rowperm_workspace_bytes(M), s));
    NAB_TRY(iota_int(s, iota.as<int>(), mn, 0));
    NAB_TRY(iota_int(s, ipiv_dev, mn, 0));
    LuCtx c{s, a, lda, M, lu_leaf_width(M), ipiv_dev, iota.as<int>(), wsg.p, wsp.p, &seq_state};
    Scratch lists;
    int leaf_seq = 0;
    c.max_leaves = ceil_div(mn, (size_t)16) + 8;
    NAB_TRY(lists.alloc(c.max_leaves * kRegListInts * sizeof(int), s));
    NAB_CUDA(cudaMemsetAsync(lists.p, 0, c.max_leaves * kRegListInts * sizeof(int), s));
    c.lists = lists.as<int>(); c.leaf_seq = &leaf_seq;
    const bool lookahead = g_lu_lookahead != 0 && mn >= 4 * LU_NB && M <= 20000;
    if (lookahead) {
        static const size_t la_leaf = [] { const char* e = getenv("NAB_LU_LEAF"); const int v = e ? atoi(e) : 64; return (size_t)((v == 16 || v == 32) ? v : 64); }();
        c.W = la_leaf;
        NAB_TRY(lu_lookahead(c, N, mn));
    } else {
        NAB_TRY(lu_rec(c, 0, mn));
    }
    if (N > mn && !lookahead) {   // wide matrix: the columns right of the square part
        double* ar = a + mn * lda;
        NAB_TRY(lu_apply_swaps(c, 0, mn, ar, N - mn));
        NAB_TRY(trsm_left(s, true, true, mn, a, 1, (ptrdiff_t)lda, nullptr, nullptr, ar, 1, (ptrdiff_t)lda, N - mn));
    }
    return NA_OK;
}

// swaps/nswaps: HOST outputs (PermutationSequence pairs).
int lu_device(cudaStream_t s, size_t M, size_t N, double* a, size_t lda, size_t* swaps, size_t* nswaps) {
    const size_t mn = std::min(M, N);
    if (nswaps) *nswaps = 0;
    if (mn == 0) return NA_OK;
    Scratch ipiv;
    NAB_TRY(ipiv.alloc(mn * sizeof(int), s));
    NAB_TRY(lu_device_async(s, M, N, a, lda, ipiv.as<int>()));
    std::vector<int> h(mn);
    NAB_CUDA(cudaMemcpyAsync(h.data(), ipiv.p, mn * sizeof(int), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    // PermutationSequence::append_permutation keeps only i != i2 (permutation_sequence.rs:84-93)
    size_t len = 0;
    if (swaps) {
        for (size_t i = 0; i < mn; ++i)
            if ((size_t)h[i] != i) { swaps[2 * len] = i; swaps[2 * len + 1] = (size_t)h[i]; ++len; }
        for (size_t i = 2 * len; i < 2 * mn; ++i) swaps[i] = 0;
    }
    if (nswaps) *nswaps = len;
    return NA_OK;
}

// LU::solve_mut: permute rows, unit-lower solve, upper solve with zero-diagonal check.
int lu_solve_device(cudaStream_t s, size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                    double* b, size_t ldb, size_t nrhs) {
    if (n == 0 || nrhs == 0) return NA_OK;
    if (n > 0x7fffff00ull) { set_error("lu_solve: dimension exceeds 2^31"); return NA_EINVAL; }
    if (nswaps) {
        std::vector<int> h(2 * nswaps);
        for (size_t i = 0; i < 2 * nswaps; ++i) {
            if (swaps[i] >= n) { set_error("lu_solve: swap index out of range"); return NA_EINVAL; }
            h[i] = (int)swaps[i];
        }
        Scratch dsw, wsp;
        NAB_TRY(dsw.alloc(2 * nswaps * sizeof(int), s));
        NAB_TRY(wsp.alloc(rowperm_workspace_bytes(n), s));
        NAB_CUDA(cudaMemcpyAsync(dsw.p, h.data(), 2 * nswaps * sizeof(int), cudaMemcpyHostToDevice, s));
        NAB_TRY(rowperm_build(s, dsw.as<int>(), dsw.as<int>() + 1, nswaps, 2, n, wsp.p));
        NAB_TRY(rowperm_apply(s, b, ldb, nrhs, std::min(2 * nswaps, n), wsp.p, n));
        NAB_CUDA(cudaStreamSynchronize(s));   // h must outlive the copy
    }
    NAB_TRY(trsm_left(s, true, true, n, lu, 1, (ptrdiff_t)lda, nullptr, nullptr, b, 1, (ptrdiff_t)ldb, nrhs));
    Scratch flag;
    NAB_TRY(flag.alloc(sizeof(int), s));
    NAB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
    NAB_TRY(zero_diag_check(s, lu, lda, nullptr, n, flag.as<int>()));
    int hflag = 0;
    NAB_CUDA(cudaMemcpyAsync(&hflag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    NAB_CUDA(cudaStreamSynchronize(s));
    if (hflag) return NA_SINGULAR;                                 // solve.rs:169-171
    return trsm_left(s, false, false, n, lu, 1, (ptrdiff_t)lda, nullptr, nullptr, b, 1, (ptrdiff_t)ldb, nrhs);
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_lu_f64_dev(size_t m, size_t n, double* a, size_t lda, size_t* swaps, size_t* nswaps, void* stream) {
    NAB_TRY(ensure_init());
    return lu_device(static_cast<cudaStream_t>(stream), m, n, a, lda, swaps, nswaps);
}

int na_lu_f64_dev_async(size_t m, size_t n, double* a, size_t lda, int32_t* ipiv_dev, void* stream) {
    NAB_TRY(ensure_init());
    if (std::min(m, n) && (!a || !ipiv_dev)) { set_error("lu (async): null pointer"); return NA_EINVAL; }
    return lu_device_async(static_cast<cudaStream_t>(stream), m, n, a, lda, ipiv_dev);
}

int na_apply_ipiv_f64_dev(size_t nrows, double* a, size_t lda, size_t ncols, const int32_t* ipiv_dev, size_t k, size_t row0,
                          void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nrows == 0 || ncols == 0 || k == 0) return NA_OK;
    if (!a || !ipiv_dev || lda < nrows || nrows > 0x7fffff00ull || row0 + k > nrows) { set_error("apply_ipiv: bad arguments"); return NA_EINVAL; }
    Scratch wsp;
    NAB_TRY(wsp.alloc(rowperm_workspace_bytes(nrows), s));
    NAB_TRY(rowperm_build_ipiv(s, ipiv_dev, k, (int)row0, nrows, wsp.p));
    return rowperm_apply(s, a, lda, ncols, std::min(2 * k, nrows), wsp.p, nrows);
}

int na_lu_f64(size_t m, size_t n, double* a, size_t lda, size_t* swaps, size_t* nswaps) {
    NAB_TRY(ensure_init());
    if (nswaps) *nswaps = 0;
    if (m == 0 || n == 0) return NA_OK;
    if (!a || lda < m) { set_error("lu: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch d; size_t ldd;
    NAB_TRY(upload_matrix(s, d, ldd, a, lda, m, n));
    NAB_TRY(lu_device(s, m, n, d.as<double>(), ldd, swaps, nswaps));
    NAB_TRY(download_matrix(s, a, lda, d.as<double>(), ldd, m, n));
    NAB_CUDA(cudaStreamSynchronize(s));
    return NA_OK;
}

int na_permute_rows_f64_dev(size_t nrows, double* a, size_t lda, size_t ncols, const size_t* swaps, size_t nswaps,
                            int inverse, void* stream) {
    NAB_TRY(ensure_init());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nrows == 0 || ncols == 0 || nswaps == 0) return NA_OK;
    if (!a || !swaps || lda < nrows || nrows > 0x7fffff00ull) { set_error("permute_rows: bad arguments"); return NA_EINVAL; }
    std::vector<int> h(2 * nswaps);
    for (size_t i = 0; i < nswaps; ++i) {
        const size_t src = inverse ? nswaps - 1 - i : i;
        if (swaps[2 * src] >= nrows || swaps[2 * src + 1] >= nrows) { set_error("permute_rows: swap index out of range"); return NA_EINVAL; }
        h[2 * i] = (int)swaps[2 * src]; h[2 * i + 1] = (int)swaps[2 * src + 1];
    }
    Scratch dsw, wsp;
    NAB_TRY(dsw.alloc(2 * nswaps * sizeof(int), s));
    NAB_TRY(wsp.alloc(rowperm_workspace_bytes(nrows), s));
    NAB_CUDA(cudaMemcpyAsync(dsw.p, h.data(), 2 * nswaps * sizeof(int), cudaMemcpyHostToDevice, s));
    NAB_TRY(rowperm_build(s, dsw.as<int>(), dsw.as<int>() + 1, nswaps, 2, nrows, wsp.p));
    NAB_TRY(rowperm_apply(s, a, lda, ncols, std::min(2 * nswaps, nrows), wsp.p, nrows));
    NAB_CUDA(cudaStreamSynchronize(s));   // h must outlive the copy
    return NA_OK;
}

int na_lu_solve_f64_dev(size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                        double* b, size_t ldb, size_t nrhs, void* stream) {
    NAB_TRY(ensure_init());
    return lu_solve_device(static_cast<cudaStream_t>(stream), n, lu, lda, swaps, nswaps, b, ldb, nrhs);
}

int na_lu_solve_f64(size_t n, const double* lu, size_t lda, const size_t* swaps, size_t nswaps,
                    double* b, size_t ldb, size_t nrhs) {
    NAB_TRY(ensure_init());
    if (n == 0 || nrhs == 0) return NA_OK;
    if (!lu || !b || lda < n || ldb < n || (nswaps && !swaps)) { set_error("lu_solve: bad arguments"); return NA_EINVAL; }
    std::lock_guard<std::mutex> lock(host_api_mutex());
    cudaStream_t s = ctx().stream;
    Scratch dl, db; size_t ldl, lddb;
    NAB_TRY(upload_matrix(s, dl, ldl, lu, lda, n, n));
    NAB_TRY(upload_matrix(s, db, lddb, b, ldb, n, nrhs));
    int st = lu_solve_device(s, n, dl.as<double>(), ldl, swaps, nswaps, db.as<double>(), lddb, nrhs);
    if (st < 0) return st;
    NAB_TRY(download_matrix(s, b, ldb, db.as<double>(), lddb, n, nrhs));   // garbage on NA_SINGULAR, like the reference
    NAB_CUDA(cudaStreamSynchronize(s));
    return st;
}

}  // extern "C"
