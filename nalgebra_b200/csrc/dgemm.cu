// dgemm.cu -- the f64 GEMM tile engine of libnalgebra_b200 (sm_100a).
//
// Replaces matrixmultiply::dgemm as called by nalgebra's gemm_uninit
// (/root/reference/src/base/blas_uninit.rs:298-313) and is the trailing-update engine of the
// blocked Cholesky / LU / QR (GEMM, SYRK-shaped lower-only GEMM).
//
// Design (B200-first):
//   * tcgen05/TMEM has no f64 kind, so the math is DMMA `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4),
//     which measures 37.18 TFLOP/s = 100 % of nominal on this part (profiles/fp64_peak_r01.md).
//   * Persistent kernel, one CTA per SM: 1 TMA producer warp + 8 DMMA consumer warps.
//     CTA tile 128x128x16, warp tile 64x32, 6-stage ring of 32 KB stages, mbarrier full/empty.
//   * Operands are staged by TMA (cp.async.bulk.tensor.2d) with SWIZZLE_128B into shared memory.
//     Either operand may be "MN-major" (unit stride along m or n) or "K-major" (unit stride along
//     k), which covers NN/NT/TN/TT without transposing anything: the m8n8k4 fragment is
//     (mn = lane/4, k = lane%4) for both A and B, and the row permutations below make the LDS.64
//     fragment reads conflict-free under the 128B swizzle for both majors.
//   * Epilogue C = alpha*acc + beta*C straight from registers; C is not read when beta == 0.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace nab {

// ------------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    });
    return fn;
}

// 2D f64 tensor map: dim0 (unit stride) has `inner` elements, dim1 has `outer` elements with a
// stride of `ld` elements.  Box = box0 x box1, SWIZZLE_128B (box0 * 8 bytes must be <= 128).
static int make_map_f64(CUtensorMap* map, const double* base, size_t inner, size_t outer, size_t ld,
                        uint32_t box0, uint32_t box1) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return NA_ECUDA; }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(double)};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%zu outer=%zu ld=%zu", (int)r, (const void*)base, inner, outer, ld);
        return NA_ECUDA;
    }
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// tile configuration
// ------------------------------------------------------------------------------------------------
namespace cfg {
constexpr int BM = 128, BN = 128, BK = 16;
constexpr int WARPS_M = 2, WARPS_N = 4;          // consumer warps
constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;   // 64 x 32
constexpr int MT = WM / 8, NT = WN / 8;          // 8 x 4 DMMA tiles per warp
constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;   // + one producer warpgroup (warp 8 issues TMA)
constexpr int STAGES = 6;
constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 2 * STAGES * 8;
constexpr int GROUP_M = 16;                      // tile rasterisation: super-rows of 16 tiles
}  // namespace cfg

// Row permutations that make the fragment reads bank-conflict free under SWIZZLE_128B.
// MN-major tile: smem = [chunk of 16 mn][k][16 mn], 16B-unit u of row k stored at u ^ (k & 7).
//   DMMA row g of tile t  <->  mn = (t/2)*16 + 4*(t%2) + perm16(g),  perm16 = {0,1,8,9,2,3,10,11}
// K-major tile: smem = [mn][16 k], 16B-unit u of row mn stored at u ^ (mn & 7).
//   DMMA row g of tile t  <->  mn = t*8 + perm8(g),                  perm8  = {0,2,4,6,1,3,5,7}
__device__ __forceinline__ int perm16(int g) { return (g & 1) | (((g >> 1) & 1) << 3) | ((g >> 2) << 1); }
__device__ __forceinline__ int perm8(int g) { return ((g & 3) << 1) | (g >> 2); }

template <bool KMAJOR>
__device__ __forceinline__ int frag_row(int t, int g) {   // row inside the warp tile
    return KMAJOR ? t * 8 + perm8(g) : (t >> 1) * 16 + ((t & 1) << 2) + perm16(g);
}
// Byte offset inside an operand tile of the element (mn = row0 + frag_row(t, g), k = 4*kk + q),
// split as  base[variant] (4 registers per operand, fixed per thread)  +  compile-time immediate,
// so every fragment read is one LDS.64 [reg + imm].
template <bool KMAJOR>
struct FragAddr {
    uint32_t base[4];
    __device__ __forceinline__ void init(int row0, int g, int q) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (KMAJOR) {          // variant = kk;  r & 7 == perm8(g) because row0 % 8 == 0
                const int r7 = perm8(g);
                base[v] = (row0 + r7) * 128 + ((((2 * v + (q >> 1)) ^ r7) << 4) | ((q & 1) << 3));
            } else {               // variant = (t & 1) * 2 + (kk & 1)
                const int mm = ((v >> 1) << 2) + perm16(g);
                const int k7 = ((v & 1) << 2) + q;
                base[v] = (row0 >> 4) * 2048 + q * 128 + ((((mm >> 1) ^ k7) << 4) | ((mm & 1) << 3));
            }
        }
    }
    __device__ __forceinline__ uint32_t off(int t, int kk) const {
        return KMAJOR ? base[kk] + t * 1024 : base[((t & 1) << 1) + (kk & 1)] + (t >> 1) * 2048 + kk * 512;
    }
};

struct GemmParams {
    int M, N, K;
    int tiles_m, tiles_n, num_tiles;
    long long ldc;
    double* C;
    double alpha, beta;
    int lower_only;   // 1: SYRK-shaped, only tiles touching the lower triangle; store row >= col only
    int splits;       // split-K: work item = (tile, split); each split writes its raw partial sums to
    int kb_per_split; //   C + split*split_stride (alpha = 1, beta = 0); splitk_reduce_kernel finishes the job
    long long split_stride;
};

// lower_only (SYRK-shaped, C is M x N with M >= N, BM == BN): t enumerates the tiles with tm >= tn,
// tile column by tile column: column tn holds T - tn tiles, prefix(tn) = tn*T - tn*(tn-1)/2, T = tiles_m.
__device__ __forceinline__ void tile_coords_lower(int t, int T, int& tm, int& tn) {
    const double b = 2.0 * T + 1.0;
    int c = (int)((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
    c = max(0, min(c, T - 1));
    while (c + 1 < T && (long long)(c + 1) * T - (long long)(c + 1) * c / 2 <= t) ++c;
    while (c > 0 && (long long)c * T - (long long)c * (c - 1) / 2 > t) --c;
    tn = c;
    tm = c + (t - (int)((long long)c * T - (long long)c * (c - 1) / 2));
}

__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
    const int group = cfg::GROUP_M * tiles_n;
    const int gid = t / group;
    const int first = gid * cfg::GROUP_M;
    const int gsz = min(tiles_m - first, cfg::GROUP_M);
    const int r = t - gid * group;
    tm = first + r % gsz;
    tn = r / gsz;
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(cfg::THREADS, 1)
dgemm_tma_dmma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const GemmParams p) {
    using namespace cfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;   // full[s] at +8*s, empty[s] at +8*(STAGES+s)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar_base + 8 * s, 1);
            ptx::mbar_init(bar_base + 8 * (STAGES + s), CONSUMER_WARPS);
        }
        ptx::fence_mbar_init();
    }
    __syncthreads();

    const int num_tiles = p.num_tiles;
    const int kblocks = (p.K + BK - 1) / BK;

    if (warp >= CONSUMER_WARPS) {
        // ================= TMA producer warpgroup (one elected thread issues) =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == CONSUMER_WARPS && lane == 0) {
            ptx::prefetch_tensormap(&map_a);
            ptx::prefetch_tensormap(&map_b);
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int tm, tn;
                const int tile = t / p.splits, split = t - tile * p.splits;
                if (p.lower_only) tile_coords_lower(tile, p.tiles_m, tm, tn); else tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
                const int kb0 = split * p.kb_per_split, kb1 = min(kblocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(bar_base + 8 * (STAGES + stage), phase ^ 1);
                    const uint32_t full = bar_base + 8 * stage;
                    const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    ptx::mbar_arrive_expect_tx(full, STAGE_BYTES);
                    if (A_KMAJOR) {
                        ptx::tma_load_2d(sa, &map_a, full, kb * BK, tm * BM);
                    } else {
#pragma unroll
                        for (int c = 0; c < BM / 16; ++c) ptx::tma_load_2d(sa + c * 2048, &map_a, full, tm * BM + c * 16, kb * BK);
                    }
                    if (B_KMAJOR) {
                        ptx::tma_load_2d(sb, &map_b, full, kb * BK, tn * BN);
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 16; ++c) ptx::tma_load_2d(sb + c * 2048, &map_b, full, tn * BN + c * 16, kb * BK);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ================= DMMA consumers (two warpgroups) =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = lane >> 2, q = lane & 3;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int row0 = wm * WM, col0 = wn * WN;

    FragAddr<A_KMAJOR> fa;
    FragAddr<B_KMAJOR> fb;
    fa.init(row0, g, q);
    fb.init(col0, g, q);

    int stage = 0;
    uint32_t phase = 0;
    // Stage hand-back.  The slot may be refilled by TMA as soon as the empty barrier completes, so every fragment
    // LDS of the stage must have RETURNED its data before this warp arrives.  (Round 1: with a plain arrive at the
    // end of the iteration ptxas sank the last k-step's DMMAs below it, that step's final LDS was still in flight
    // when the slot was handed back, and under a backed-up LSU the refill won the race about once per 1e8 stage
    // reads.)  Two independent guards:
    //   1. an explicit DATA dependency: `frag_dep` folds the high words of all 12 fragment registers of the stage's
    //      last k-step into a value that is always 0 but opaque to the compiler (x*x + x is even), and the arrive
    //      adds it to its count operand -- the arrive cannot issue before those loads have written their registers
    //      (shared-memory loads of a warp return in order, so the earlier k-steps' loads are covered too);
    //   2. the arrive is deferred until after the NEXT stage's full-barrier wait, as before.
    // Gating regression: tools/gemm_stress.py (repeated full-size beta != 0 launches must agree bit for bit) and
    // test_gemm_full_size_linearity_on_device; validated with nvcc 12.9.86.
    int release_stage = -1;
    uint32_t release_dep = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int tm, tn;
        const int tile = t / p.splits, split = t - tile * p.splits;
        if (p.lower_only) tile_coords_lower(tile, p.tiles_m, tm, tn); else tile_coords(tile, p.tiles_m, p.tiles_n, tm, tn);
        const int kb0 = split * p.kb_per_split, kb1 = min(kblocks, kb0 + p.kb_per_split);

        double acc[MT][NT][2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

        for (int kb = kb0; kb < kb1; ++kb) {
            ptx::mbar_wait(bar_base + 8 * stage, phase);
            if (release_stage >= 0 && lane == 0) ptx::mbar_arrive_cnt(bar_base + 8 * (STAGES + release_stage), 1u + release_dep);
            const uint8_t* sptr = smem_gen + stage * STAGE_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                double a[MT], b[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) a[i] = *reinterpret_cast<const double*>(sptr + fa.off(i, kk));
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = *reinterpret_cast<const double*>(sptr + A_BYTES + fb.off(j, kk));
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) ptx::dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                if (kk == 3) {
                    uint32_t x = 0;
#pragma unroll
                    for (int i = 0; i < MT; ++i) x ^= (uint32_t)__double2hiint(a[i]);
#pragma unroll
                    for (int j = 0; j < NT; ++j) x ^= (uint32_t)__double2hiint(b[j]);
                    release_dep = ptx::opaque_zero(x);
                }
            }
            release_stage = stage;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }

        // ---- epilogue: C = alpha*acc + beta*C (C untouched-before-write when beta == 0) ----
        // The C reads are issued in batches of 32 independent loads BEFORE any store of the batch: with
        // load/store interleaved per element every load waits for the previous store (possible aliasing)
        // and the beta != 0 epilogue cost ~14 us per tile instead of ~1.
        const long long gm0 = (long long)tm * BM + row0, gn0 = (long long)tn * BN + col0;
        const bool use_beta = p.beta != 0.0;
        double* cbase = p.C + (long long)split * p.split_stride;
        long long rows[MT];
        bool rok[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) { rows[i] = gm0 + frag_row<A_KMAJOR>(i, g); rok[i] = rows[i] < p.M; }
#pragma unroll
        for (int jh = 0; jh < NT; jh += 2) {
            double cv[2][2][MT];
            long long cols[2][2];
            bool cok[2][2];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    cols[jj][e] = gn0 + frag_row<B_KMAJOR>(jh + jj, 2 * q + e);
                    cok[jj][e] = cols[jj][e] < p.N;
                }
            if (use_beta) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const bool ok = cok[jj][e] && rok[i] && !(p.lower_only && rows[i] < cols[jj][e]);
                            cv[jj][e][i] = ok ? cbase[rows[i] + cols[jj][e] * p.ldc] : 0.0;
                        }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const bool ok = cok[jj][e] && rok[i] && !(p.lower_only && rows[i] < cols[jj][e]);
                        double v = p.alpha * acc[i][jh + jj][e];
                        if (use_beta) v += p.beta * cv[jj][e][i];
                        if (ok) cbase[rows[i] + cols[jj][e] * p.ldc] = v;
                    }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// auxiliary kernels: strided pack / scatter, scale, fill
// ------------------------------------------------------------------------------------------------
__global__ void pack_strided_kernel(double* __restrict__ dst, long long ldd, const double* __restrict__ src,
                                    long long rs, long long cs, long long rows, long long cols) {
    // dst is column-major rows x cols (ldd).  Tile through shared memory so that both sides are
    // coalesced whichever of rs/cs is the small stride.
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
    const bool src_row_fast = (rs < 0 ? -rs : rs) <= (cs < 0 ? -cs : cs);
    if (src_row_fast) {
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            long long r = r0 + threadIdx.x, c = c0 + j;
            if (r < rows && c < cols) dst[r + c * ldd] = src[r * rs + c * cs];
        }
        return;
    }
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        long long r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * rs + c * cs];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long long r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) dst[r + c * ldd] = tile[threadIdx.x][j];
    }
}

__global__ void scatter_strided_kernel(double* __restrict__ dst, long long rs, long long cs,
                                       const double* __restrict__ src, long long lds, long long rows, long long cols) {
    __shared__ double tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        long long r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[threadIdx.x][j] = src[r + c * lds];
    }
    __syncthreads();
    const bool dst_row_fast = (rs < 0 ? -rs : rs) <= (cs < 0 ? -cs : cs);
    if (dst_row_fast) {
        for (int j = threadIdx.y; j < 32; j += blockDim.y) {
            long long r = r0 + threadIdx.x, c = c0 + j;
            if (r < rows && c < cols) dst[r * rs + c * cs] = tile[threadIdx.x][j];
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            long long r = r0 + i, c = c0 + threadIdx.x;
            if (r < rows && c < cols) dst[r * rs + c * cs] = tile[i][threadIdx.x];
        }
    }
}

// C <- beta*C (beta == 0 writes zeros without reading C): the k == 0 contract of
// gemm_uninit, /root/reference/src/base/blas_uninit.rs:258-269 and :152-160.
__global__ void scale_strided_kernel(double* __restrict__ c, long long rs, long long cs, long long rows, long long cols,
                                     double beta) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % rows, col = idx / rows;
        double* p = c + r * rs + col * cs;
        *p = (beta == 0.0) ? 0.0 : (*p * beta);
    }
}

// lower triangle (incl. diagonal) of a w x w block: C <- beta * C, zeros when beta == 0 (C not read)
__global__ void scale_lower_block_kernel(double* __restrict__ c, long long ldc, int w, double beta) {
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int r = idx % w, col = idx / w;
        if (r >= col) { double* p = c + r + col * ldc; *p = (beta == 0.0) ? 0.0 : (*p * beta); }
    }
}
int scale_lower_block(cudaStream_t s, double* c, size_t ldc, size_t w, double beta) {
    if (w == 0) return NA_OK;
    scale_lower_block_kernel<<<1, 256, 0, s>>>(c, (long long)ldc, (int)w, beta);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
// a(i,j) = rand01(seed, (row0+i) + (col0+j)*global_rows): a block of the global matrix.
__global__ void fill_uniform_kernel(double* __restrict__ a, long long nrows, long long ncols, long long lda, uint64_t seed,
                                    long long row0, long long col0, long long global_rows) {
    const long long total = nrows * ncols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % nrows, c = idx / nrows;
        const uint64_t gidx = (uint64_t)((row0 + r) + (col0 + c) * global_rows);
        const uint64_t h = mix64(gidx + (seed + 1) * 0x9E3779B97F4A7C15ULL);
        a[r + c * lda] = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    }
}

__global__ void copy_strided_kernel(double* __restrict__ dst, long long rsd, long long csd, const double* __restrict__ src,
                                    long long rss, long long css, long long rows, long long cols, int dst_row_fast) {
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r, c;
        if (dst_row_fast) { r = idx % rows; c = idx / rows; } else { c = idx % cols; r = idx / cols; }
        dst[r * rsd + c * csd] = src[r * rss + c * css];
    }
}
int copy_strided(cudaStream_t s, double* dst, ptrdiff_t rsd, ptrdiff_t csd, const double* src, ptrdiff_t rss, ptrdiff_t css,
                 size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return NA_OK;
    const int row_fast = (rsd < 0 ? -rsd : rsd) <= (csd < 0 ? -csd : csd) ? 1 : 0;
    int blocks = (int)std::min<size_t>(ceil_div(rows * cols, 256), (size_t)ctx().sm_count * 16);
    copy_strided_kernel<<<blocks, 256, 0, s>>>(dst, rsd, csd, src, rss, css, (long long)rows, (long long)cols, row_fast);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// (B + B^T)/2 + n*I for the block [row0,+nrows) x [col0,+ncols), B(i,j) = rand01(seed, i + j*n)
__global__ void fill_spd_kernel(double* __restrict__ a, long long nrows, long long ncols, long long lda, uint64_t seed,
                                long long row0, long long col0, long long n) {
    const long long total = nrows * ncols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % nrows, c = idx / nrows, gi = row0 + r, gj = col0 + c;
        const uint64_t k = (seed + 1) * 0x9E3779B97F4A7C15ULL;
        const double bij = (double)(mix64((uint64_t)(gi + gj * n) + k) >> 11) * (1.0 / 9007199254740992.0);
        const double bji = (double)(mix64((uint64_t)(gj + gi * n) + k) >> 11) * (1.0 / 9007199254740992.0);
        a[r + c * lda] = (bij + bji) * 0.5 + (gi == gj ? (double)n : 0.0);
    }
}
int fill_spd(cudaStream_t s, double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed, size_t row0, size_t col0, size_t n) {
    if (nrows == 0 || ncols == 0) return NA_OK;
    int blocks = (int)std::min<size_t>(ceil_div(nrows * ncols, 256), (size_t)ctx().sm_count * 16);
    fill_spd_kernel<<<blocks, 256, 0, s>>>(a, (long long)nrows, (long long)ncols, (long long)lda, seed, (long long)row0, (long long)col0, (long long)n);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

int pack_strided(cudaStream_t s, double* dst, size_t ldd, const double* src, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return NA_OK;
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32)), block(32, 8);
    if (grid.y > 65535) { set_error("pack_strided: too many columns (%zu)", cols); return NA_EINVAL; }
    pack_strided_kernel<<<grid, block, 0, s>>>(dst, (long long)ldd, src, rs, cs, (long long)rows, (long long)cols);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}
int scatter_strided(cudaStream_t s, double* dst, ptrdiff_t rs, ptrdiff_t cs, const double* src, size_t lds, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return NA_OK;
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32)), block(32, 8);
    if (grid.y > 65535) { set_error("scatter_strided: too many columns (%zu)", cols); return NA_EINVAL; }
    scatter_strided_kernel<<<grid, block, 0, s>>>(dst, rs, cs, src, (long long)lds, (long long)rows, (long long)cols);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}
int scale_strided(cudaStream_t s, double* c, ptrdiff_t rs, ptrdiff_t cs, size_t rows, size_t cols, double beta) {
    if (rows == 0 || cols == 0) return NA_OK;
    size_t total = rows * cols;
    int blocks = (int)std::min<size_t>(ceil_div(total, 256), (size_t)ctx().sm_count * 8);
    scale_strided_kernel<<<blocks, 256, 0, s>>>(c, rs, cs, (long long)rows, (long long)cols, beta);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}
int fill_uniform(cudaStream_t s, double* a, size_t nrows, size_t ncols, size_t lda, uint64_t seed,
                 size_t row0, size_t col0, size_t global_rows) {
    if (nrows == 0 || ncols == 0) return NA_OK;
    size_t total = nrows * ncols;
    int blocks = (int)std::min<size_t>(ceil_div(total, 256), (size_t)ctx().sm_count * 16);
    fill_uniform_kernel<<<blocks, 256, 0, s>>>(a, (long long)nrows, (long long)ncols, (long long)lda, seed,
                                               (long long)row0, (long long)col0, (long long)global_rows);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
struct Operand {          // logical [mn x k] operand
    const double* ptr;
    ptrdiff_t s_mn, s_k;  // element strides
    size_t mn, k;
};

static bool tma_ok(const double* base, size_t ld) {
    return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld % 2) == 0 && ld > 0;
}

// Chooses the TMA view of an operand, packing it into `scratch` first when its layout cannot be
// described by a tensor map (both strides != 1, negative strides, or 16-byte misalignment).
static int prepare_operand(cudaStream_t s, const Operand& op, Scratch& scratch, bool& kmajor, const double*& base, size_t& ld) {
    if (op.s_mn == 1 && op.s_k >= (ptrdiff_t)op.mn && tma_ok(op.ptr, (size_t)op.s_k)) {
        kmajor = false; base = op.ptr; ld = (size_t)op.s_k; return NA_OK;
    }
    if (op.s_k == 1 && op.s_mn >= (ptrdiff_t)op.k && tma_ok(op.ptr, (size_t)op.s_mn)) {
        kmajor = true; base = op.ptr; ld = (size_t)op.s_mn; return NA_OK;
    }
    // degenerate single row/column views whose other stride is irrelevant
    if (op.s_mn == 1 && op.k == 1 && tma_ok(op.ptr, round_up(op.mn, 2))) {
        kmajor = false; base = op.ptr; ld = round_up(op.mn, 2); return NA_OK;
    }
    // pack to MN-major, ld even
    ld = round_up(op.mn, 2);
    NAB_TRY(scratch.alloc(ld * op.k * sizeof(double), s));
    NAB_TRY(pack_strided(s, scratch.as<double>(), ld, op.ptr, op.s_mn, op.s_k, op.mn, op.k));
    kmajor = false; base = scratch.as<double>();
    return NA_OK;
}

// Upper bound on the CTAs a GEMM launch may use (look-ahead: panel work and the bulk trailing update
// run concurrently on disjoint sets of SMs).  0 = all SMs.  Thread local: set by the blocked drivers.
// Two levels: the caller's own reservation (na_set_gemm_sm_limit) and the limit a blocked driver sets around its panel /
// bulk launches; the smaller non-zero one applies, and a driver clearing ITS limit leaves the caller's in place.
static thread_local int g_sm_limit = 0, g_user_sm_limit = 0;
void set_gemm_sm_limit(int limit) { g_sm_limit = limit; }
void set_user_gemm_sm_limit(int limit) { g_user_sm_limit = limit; }
static int effective_sm_limit() {
    if (g_sm_limit > 0 && g_user_sm_limit > 0) return std::min(g_sm_limit, g_user_sm_limit);
    return g_sm_limit > 0 ? g_sm_limit : g_user_sm_limit;
}

static int launch_gemm(cudaStream_t s, bool a_kmajor, bool b_kmajor, const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p) {
    using namespace cfg;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        cudaError_t e;
        e = cudaFuncSetAttribute(dgemm_tma_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e) attr_err = e;
        e = cudaFuncSetAttribute(dgemm_tma_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e) attr_err = e;
        e = cudaFuncSetAttribute(dgemm_tma_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e) attr_err = e;
        e = cudaFuncSetAttribute(dgemm_tma_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); if (e) attr_err = e;
    });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(dgemm)", __FILE__, __LINE__);
    int grid = std::min(p.num_tiles, ctx().sm_count);
    if (effective_sm_limit() > 0) grid = std::min(grid, effective_sm_limit());
    if (a_kmajor) {
        if (b_kmajor) dgemm_tma_dmma_kernel<true, true><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, p);
        else dgemm_tma_dmma_kernel<true, false><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, p);
    } else {
        if (b_kmajor) dgemm_tma_dmma_kernel<false, true><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, p);
        else dgemm_tma_dmma_kernel<false, false><<<grid, THREADS, SMEM_BYTES, s>>>(ma, mb, p);
    }
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

// C <- alpha * sum_s ws[s] + beta * C   (split-K epilogue; C not read when beta == 0)
__global__ void splitk_reduce_kernel(double* __restrict__ c, long long ldc, const double* __restrict__ ws, long long ldw,
                                     long long split_stride, int splits, long long m, long long n, double alpha, double beta) {
    const long long total = m * n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx % m, col = idx / m;
        // partial sums are added in split order (deterministic); 8 independent loads in flight per thread --
        // one load per iteration made this kernel latency-bound (50 us for 148 splits of a 32 x 224 block)
        const double* w0 = ws + r + col * ldw;
        double s = 0.0;
        int k = 0;
        for (; k + 8 <= splits; k += 8) {
            double v8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v8[i] = w0[(long long)(k + i) * split_stride];
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v8[i];
        }
        for (; k < splits; ++k) s += w0[(long long)k * split_stride];
        double v = alpha * s;
        if (beta != 0.0) v += beta * c[r + col * ldc];
        c[r + col * ldc] = v;
    }
}

// C(m x n, unit row stride, ldc) <- alpha * A * B + beta * C on device.
static int gemm_colmajor_c(cudaStream_t s, bool lower_only, size_t m, size_t n, size_t k, double alpha,
                           const Operand& A, const Operand& B, double beta, double* c, size_t ldc) {
    using namespace cfg;
    if (m > 0x7fffff00ull || n > 0x7fffff00ull || k > 0x7fffff00ull) { set_error("gemm: dimension exceeds 2^31"); return NA_EINVAL; }
    Scratch sa, sb;
    bool a_km, b_km;
    const double *abase, *bbase;
    size_t lda, ldb;
    NAB_TRY(prepare_operand(s, A, sa, a_km, abase, lda));
    NAB_TRY(prepare_operand(s, B, sb, b_km, bbase, ldb));
    CUtensorMap ma, mb;
    if (a_km) NAB_TRY(make_map_f64(&ma, abase, k, m, lda, BK, BM)); else NAB_TRY(make_map_f64(&ma, abase, m, k, lda, 16, BK));
    if (b_km) NAB_TRY(make_map_f64(&mb, bbase, k, n, ldb, BK, BN)); else NAB_TRY(make_map_f64(&mb, bbase, n, k, ldb, 16, BK));
    GemmParams p;
    p.M = (int)m; p.N = (int)n; p.K = (int)k;
    p.tiles_m = (int)ceil_div(m, BM); p.tiles_n = (int)ceil_div(n, BN);
    p.ldc = (long long)ldc; p.C = c; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only ? 1 : 0;
    p.splits = 1; p.kb_per_split = (int)ceil_div(k, BK); p.split_stride = 0;
    if (lower_only) {
        if (m < n) { set_error("gemm: lower_only needs C with rows >= cols"); return NA_EINVAL; }
        const long long T = p.tiles_m, Tn = p.tiles_n;
        p.num_tiles = (int)(Tn * T - Tn * (Tn - 1) / 2);
    } else {
        if ((long long)p.tiles_m * p.tiles_n > 0x7fffffffLL) { set_error("gemm: too many tiles"); return NA_EINVAL; }
        p.num_tiles = p.tiles_m * p.tiles_n;
        // split-K for few-tile / deep-K shapes (V^T C in the blocked QR, Gram matrices)
        const int sms = ctx().sm_count;
        const int kblocks = (int)ceil_div(k, BK);
        auto wave_eff = [&](int sp) {
            const long long items = (long long)p.num_tiles * sp;
            return (double)items / (double)(ceil_div((size_t)items, (size_t)sms) * (size_t)sms);
        };
        if (kblocks >= 64 && wave_eff(1) < 0.8) {
            // K-slices chosen for wave efficiency over the persistent grid (work items = tiles x slices are dealt
            // round-robin to the CTAs): the fewest slices that fill >= 93 % of the last wave, else the best found;
            // at least 16 k-blocks (256 deep) per slice.  (60 tiles: 2 slices = 81 %, 7 slices = 95 %.)
            const int smax = std::min(kblocks / 16, 148);
            int splits = 1;
            double best = wave_eff(1);
            for (int sp = 2; sp <= smax; ++sp) {
                const double e = wave_eff(sp);
                if (e > best + 1e-9) { best = e; splits = sp; }
                if (best >= 0.93) break;
            }
            while (splits > 1 && (size_t)splits * round_up(m, 2) * n * sizeof(double) > (512ull << 20)) --splits;
            if (splits > 1) {
                const int kbs = (int)ceil_div((size_t)kblocks, (size_t)splits);
                splits = (int)ceil_div((size_t)kblocks, (size_t)kbs);
                Scratch ws;
                const size_t ldw = round_up(m, 2);
                NAB_TRY(ws.alloc((size_t)splits * ldw * n * sizeof(double), s));
                GemmParams ps = p;
                ps.C = ws.as<double>(); ps.ldc = (long long)ldw; ps.alpha = 1.0; ps.beta = 0.0;
                ps.splits = splits; ps.kb_per_split = kbs; ps.split_stride = (long long)(ldw * n);
                ps.num_tiles = p.num_tiles * splits;
                NAB_TRY(launch_gemm(s, a_km, b_km, ma, mb, ps));
                const int blocks = (int)std::min<size_t>(ceil_div(m * n, 256), (size_t)sms * 8);
                splitk_reduce_kernel<<<blocks, 256, 0, s>>>(c, (long long)ldc, ws.as<double>(), (long long)ldw, ps.split_stride, splits,
                                                            (long long)m, (long long)n, alpha, beta);
                NAB_LAUNCH_CHECK();
                return NA_OK;
            }
        }
    }
    return launch_gemm(s, a_km, b_km, ma, mb, p);
}

int dgemm_device(cudaStream_t s, bool lower_only, size_t m, size_t k, size_t n, double alpha,
                 const double* a, ptrdiff_t rsa, ptrdiff_t csa, const double* b, ptrdiff_t rsb, ptrdiff_t csb,
                 double beta, double* c, ptrdiff_t rsc, ptrdiff_t csc) {
    NAB_TRY(ensure_init());
    if (m == 0 || n == 0) return NA_OK;
    if (k == 0) return scale_strided(s, c, rsc, csc, m, n, beta);   // blas_uninit.rs:258-269
    if (!a || !b || !c) { set_error("gemm: null pointer"); return NA_EINVAL; }
    Operand A{a, rsa, csa, m, k};      // [m x k]: s_mn = rsa, s_k = csa
    Operand B{b, csb, rsb, n, k};      // [n x k]: s_mn = csb, s_k = rsb
    if (rsc == 1 && (csc >= (ptrdiff_t)m || n == 1)) return gemm_colmajor_c(s, lower_only, m, n, k, alpha, A, B, beta, c, (size_t)csc);
    if (csc == 1 && (rsc >= (ptrdiff_t)n || m == 1) && !lower_only)   // row-major C: C^T = B^T A^T
        return gemm_colmajor_c(s, false, n, m, k, alpha, B, A, beta, c, (size_t)rsc);
    // general C strides: compute into a packed temporary and scatter
    if (lower_only) { set_error("gemm: lower_only needs a column-major C"); return NA_EINVAL; }
    Scratch tmp;
    size_t ldt = round_up(m, 2);
    NAB_TRY(tmp.alloc(ldt * n * sizeof(double), s));
    if (beta != 0.0) NAB_TRY(pack_strided(s, tmp.as<double>(), ldt, c, rsc, csc, m, n));
    NAB_TRY(gemm_colmajor_c(s, false, m, n, k, alpha, A, B, beta, tmp.as<double>(), ldt));
    return scatter_strided(s, c, rsc, csc, tmp.as<double>(), ldt, m, n);
}

}  // namespace nab
