// kernels.cuh -- internal C++ interface between the translation units of libnalgebra_b200.
#pragma once
#include "common.cuh"

namespace nab {

// ---- dgemm.cu ------------------------------------------------------------------------------------
// C <- alpha*A*B + beta*C on device pointers with arbitrary element strides (matrixmultiply::dgemm
// semantics).  lower_only: SYRK-shaped update, only elements with row >= col of a column-major C
// are computed/stored.
int dgemm_device(cudaStream_t s, bool lower_only, size_t m, size_t k, size_t n, double alpha,
                 const double* a, ptrdiff_t rsa, ptrdiff_t csa, const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                 double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc);
int fill_spd(cudaStream_t s, double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed, size_t row0, size_t col0, size_t n);
void set_gemm_sm_limit(int limit);        // the blocked drivers' limit; 0 = none; thread local
void set_user_gemm_sm_limit(int limit);   // the caller's reservation (na_set_gemm_sm_limit); thread local
// Per-SM throughput the look-ahead schedule models assume for K = nb update GEMMs (36.3 TFLOP/s / 148 SMs at ~90 %).
constexpr double kSmFlops = 0.22e12;
int pack_strided(cudaStream_t s, double* dst, size_t ldd, const double* src, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols);
int scatter_strided(cudaStream_t s, double* dst, ptrdiff_t rs, ptrdiff_t cs, const double* src, size_t lds, size_t rows, size_t cols);
int scale_lower_block(cudaStream_t s, double* c, size_t ldc, size_t w, double beta);
int scale_strided(cudaStream_t s, double* c, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols, double beta);
int fill_uniform(cudaStream_t s, double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                 size_t row0, size_t col0, size_t global_rows);

// general strided 2D copy dst(i,j) = src(i,j)
int copy_strided(cudaStream_t s, double* dst, ptrdiff_t rsd, ptrdiff_t csd, const double* src, ptrdiff_t rss, ptrdiff_t css,
                 size_t rows, size_t cols);

// ---- sgemm_tc.cu: C (column-major m x n) <- alpha*A*B + beta*C, tcgen05 kind::tf32 3xTF32; any operand strides
int sgemm_tc_colmajor(cudaStream_t s, size_t m, size_t k, size_t n, float alpha, const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                      const float* b, ptrdiff_t rsb, ptrdiff_t csb, float beta, float* c, size_t ldc);

// ---- panel_chol_tri.cu ---------------------------------------------------------------------------
constexpr int kInvBlock = 128;     // diagonal-block size of POTF2 / TRTRI / the TRSM base case
int potf2(cudaStream_t st, double* a, size_t lda, int n, int use_sub, double sub, size_t col0, unsigned long long* fail_col,
          double* inv_out);
int trtri_blocks(cudaStream_t st, const double* t, ptrdiff_t rs, ptrdiff_t cs, size_t n, bool eff_lower, bool unit,
                 const double* diag_abs, double* out);
int zero_diag_check(cudaStream_t st, const double* t, size_t ldt, const double* diag_abs, size_t n, int* flag);
int set_identity(cudaStream_t st, double* a, size_t lda, size_t rows, size_t cols);

// ---- panel_lu.cu ---------------------------------------------------------------------------------
constexpr int kLuPanel = 128;      // widest GETF2 leaf panel
size_t getf2_workspace_bytes();
int getf2_panel(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, size_t j0, int* ipiv, void* ws, int* seq_state,
                int cta_limit = 0);
// register-resident leaf (panel_lu_reg.cu): w <= 64, rows in registers, implicit pivoting
int getf2_reg_grid(size_t m, size_t w);     // CTAs it needs for an m x w panel; 0 = does not fit
// list (optional, device, zeroed count): the rows the leaf moved, ready for rowperm_apply_lists -- [0] = count,
// [1 ..] dest[kRegListMax], then src[kRegListMax] (global row indices j0 + ...)
constexpr int kRegListMax = 128;
constexpr int kRegListInts = 1 + 2 * kRegListMax + 3;     // one leaf's list, padded to a multiple of 4 ints
int getf2_panel_reg(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, size_t j0, int* ipiv, void* ws, int* seq_state,
                    int cta_limit = 0, int* list = nullptr);
int rowperm_apply_lists(cudaStream_t st, double* a, size_t lda, size_t ncols, size_t max_touched, const int* count, const int* dest,
                        const int* src);
size_t rowperm_workspace_bytes(size_t n);
int rowperm_build(cudaStream_t st, const int* sa, const int* sb, size_t K, size_t stride, size_t n, void* ws);
int rowperm_build_ipiv(cudaStream_t st, const int* ipiv, size_t K, int row0, size_t n, void* ws);
int rowperm_apply(cudaStream_t st, double* a, size_t lda, size_t ncols, size_t max_touched, const void* ws, size_t n);
int iota_int(cudaStream_t st, int* p, size_t n, int offset);
// B (n1 x nrhs, column-major) <- L^-1 B, L = unit lower triangle of l (n1 <= 128); one launch, no inverse blocks
int trsm_unit_lower_small(cudaStream_t st, size_t n1, const double* l, size_t ldl, double* b, size_t ldb, size_t nrhs);

// ---- panel_update.cu: C -= A * B for a tall C and K = 16 / 32 / 64 (bandwidth-bound; DMMA fragments from global memory).
// Returns 1 (nothing done) for other K; max_ctas = SMs it may occupy (0 = all)
int rank_update_small_k(cudaStream_t st, size_t m, size_t k, size_t n, const double* a, size_t lda, const double* b, size_t ldb, double* c,
                        size_t ldc, int max_ctas);

// ---- panel_qr.cu ---------------------------------------------------------------------------------
constexpr int kQrLeaf = 32;        // GEQR2 leaf panel width
size_t geqr2_workspace_bytes();
int geqr2_grid(size_t m, size_t w);       // CTAs (= SMs) the cooperative GEQR2 launch of an m x w panel occupies
int geqr2_panel(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, double* tau, void* ws, int* seq_state);
// register-resident leaf (panel_qr_reg.cu): 512 rows per CTA; grid 0 = the panel does not fit the SMs
int geqr2_reg_grid(size_t m);
int geqr2_panel_reg(cudaStream_t st, double* a_panel, size_t lda, size_t m, size_t w, double* tau, void* ws, int* seq_state);
int extract_v(cudaStream_t st, double* vw, size_t ldv, const double* a, size_t lda, size_t m, size_t w, const double* tau, int mode);
int build_s(cudaStream_t st, double* g, size_t ldg, size_t w, const double* tau);
// leaf panels (w <= 32): clean V into vw and S = triu(V^T V, 1) + diag(1/tau) into smat, one launch
size_t extract_v_gram_workspace_bytes();
int extract_v_gram(cudaStream_t st, double* vw, size_t ldv, const double* a, size_t lda, size_t m, size_t w, const double* tau,
                   double* smat, size_t lds, void* ws);
// panel_qr_fused.cu: C <- (I - V T^T V^T) C for a leaf (reflector vectors in columns [0, w) of a_leaf, C right of them), one launch
size_t larfb_fused_workspace_bytes();
int larfb_leaf_fused(cudaStream_t st, double* a_leaf, size_t lda, size_t ml, size_t w, size_t nc, const double* tau, void* ws, int* seq,
                     int max_ctas);
int tau_from_diag(cudaStream_t st, double* tau_out, const double* diag, size_t n);
int qr_convert_to_nalgebra(cudaStream_t st, double* a, size_t lda, size_t m, size_t n, const double* tau, double* csign, double* diag);
int qr_convert_columns(cudaStream_t st, double* a, size_t lda, size_t m, size_t k, size_t j0, size_t ncols, const double* tau, double* csign,
                       double* diag);
int qr_signs_from_diag(cudaStream_t st, const double* diag, size_t k, double* csign);
int scale_signs(cudaStream_t st, double* b, size_t ldb, size_t rows, size_t cols, const double* csign, size_t k, bool by_cols);

// ---- factor.cu -----------------------------------------------------------------------------------
// Solves M X = B in place on B (n x nrhs, strides rsb/csb, one of them 1).  M (strides rsm/csm) is
// effectively lower or upper triangular.  inv_blocks: optional precomputed inverses of M's 128x128
// diagonal blocks (trtri_blocks layout); computed internally when null.
int trsm_left(cudaStream_t s, bool eff_lower, bool unit, size_t n, const double* m, ptrdiff_t rsm, ptrdiff_t csm,
              const double* diag_abs, const double* inv_blocks, double* b, ptrdiff_t rsb, ptrdiff_t csb, size_t nrhs);
void lu_set_lookahead(long v);
void qr_set_tuning(int which, long v);      // 0: register-resident GEQR2 leaf, 1: fused in-panel block reflector
int cholesky_device(cudaStream_t s, size_t n, double* a, size_t lda, int use_sub, double sub, size_t* fail_col);
int cholesky_solve_device(cudaStream_t s, size_t n, const double* l, size_t lda, double* b, size_t ldb, size_t nrhs);
int lu_device(cudaStream_t s, size_t M, size_t N, double* a, size_t lda, size_t* swaps, size_t* nswaps);
int lu_device_async(cudaStream_t s, size_t M, size_t N, double* a, size_t lda, int* ipiv_dev);
// sink (optional): host matrix that receives every column block as soon as it is final (host-pointer entry point)
struct QrHostSink { double* h; size_t ldh; };
int qr_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diag, double* tau_out, const QrHostSink* sink = nullptr);
int apply_q_blocks(cudaStream_t s, size_t m, size_t k, const double* qr, size_t lda, const double* diag,
                   double* b, size_t ldb, size_t nb, bool forward, bool triangular_q, const double* lapack_tau);
// factor_twosided.cu: d / e DEVICE vectors of signed norms
void ts_set_fused(long v);           // -1: size rule, 0: two-pass kernels, 1: fused one-pass kernels
int hessenberg_device(cudaStream_t s, size_t n, double* a, size_t lda, double* subdiag);
int symmetric_tridiagonal_device(cudaStream_t s, size_t n, double* a, size_t lda, double* off_diagonal);
int bidiagonal_device(cudaStream_t s, size_t m, size_t n, double* a, size_t lda, double* diagonal, double* off_diagonal);
int upload_matrix(cudaStream_t s, Scratch& buf, size_t& ldd, const double* h, size_t ldh, size_t rows, size_t cols);
int download_matrix(cudaStream_t s, double* h, size_t ldh, const double* d, size_t ldd, size_t rows, size_t cols);

}  // namespace nab
