// sgemm_tc.cu -- f32 GEMM on the 5th-generation tensor cores: TMA-staged tcgen05.mma kind::tf32 with the
// accumulators in TMEM, 3xTF32 operand splitting so that the result stays within f32 tolerance.
//
// Replaces matrixmultiply::sgemm as called by gemm_uninit (/root/reference/src/base/blas_uninit.rs:276-291).
//
// 3xTF32.  The tensor core reads an f32 operand as TF32 (10 explicit mantissa bits, the low 13 are ignored).  With
//   x_hi = x with the low 13 mantissa bits cleared,   x_lo = x - x_hi   (exact in f32)
// the product a*b is recovered to ~2^-21 relative as  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  (a_lo*b_lo ~ 2^-22 |ab| is
// dropped), accumulated in f32 inside TMEM.  Three tcgen05.mma per k-step instead of one.
//
// Data flow (B200-first):
//   1. pack: one HBM-bound pass per operand reads the caller's matrix through ARBITRARY element strides (views,
//      transposes, negative strides) and writes the hi and lo parts as dense K-major arrays [rows][kpad] (kpad = k
//      rounded up to 32, zero filled) -- the one layout in which a tile is a single TMA box and a canonical UMMA
//      K-major SWIZZLE_128B operand.  12 bytes of traffic per element against 2*k flops: ~1.5 % of the GEMM at 16384.
//   2. persistent kernel, one CTA per SM, 6 warps: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one
//      thread; the warp also owns the TMEM allocation), warps 2-5 = epilogue (TMEM -> registers -> C).
//      CTA tile 128 x 128 x 32; a stage holds A_hi, A_lo, B_hi, B_lo (4 x 16 KB); 3-stage ring, mbarrier full/empty;
//      12 tcgen05.mma (M128 N128 K8) per stage; tcgen05.commit hands the stage back and signals the epilogue.
//      Two 128-column accumulators in TMEM (256 of 512 columns): the epilogue of tile i overlaps the MMAs of i+1.
//   3. epilogue: tcgen05.ld 32x32b (lane = row, 32 columns per load), C = alpha*acc + beta*C written straight to the
//      caller's column-major C (a warp writes 32 consecutive rows of one column: coalesced); C is not read when beta = 0.
// Accumulation accuracy.  The tensor core adds into its f32 accumulator with truncation, so the error of one long TMEM
// accumulation grows linearly with k (measured 2.7e-4 relative at k = 8192).  The K loop is therefore cut into chunks of
// KC = 8 k-blocks (256 columns of k): every chunk is accumulated in TMEM, handed to the epilogue warps through the
// double-buffered accumulator, and summed there in registers with round-to-nearest f32 adds.
// Shared-memory-bandwidth bound by construction (an SS-mode 128x128x8 tf32 MMA reads 8 KB of operands for 64 cycles of
// math at 128 B/clk), so the ceiling is about half the TF32 peak, i.e. ~1/6 of it in f32-equivalent flops.
#include "common.cuh"
#include "ptx.cuh"
#include "kernels.cuh"

namespace nab {

namespace tcfg {
constexpr int BM = 128, BN = 128, BK = 32;           // BK * 4 bytes = 128 B = one swizzle row
constexpr int UMMA_K = 8;                            // tf32: 32 bytes per MMA k-step
constexpr int STAGES = 3;
constexpr int KC = 8;                                // k-blocks per TMEM accumulation chunk (see "Accumulation accuracy")
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB, the same for A and B tiles (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr int THREADS = 6 * 32;
constexpr int ACC_COLS = BN;                         // TMEM columns per accumulator
constexpr int TMEM_COLS = 2 * ACC_COLS;              // double-buffered
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers + tmem ptr*/;
}  // namespace tcfg

// ---- tcgen05 wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {     // arrives on the mbarrier when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, both operands through shared-memory descriptors
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_desc_kmajor_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = f32, A = B = TF32, both K-major, M = 128, N = BN
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(tcfg::BN >> 3) << 17) | ((uint32_t)(tcfg::BM >> 4) << 24);

// Tile rasterisation: super-rows of 16 tile rows, column by column inside a super-row, so that a wave of 148 tiles shares
// ~16 A row panels and ~9 B column panels through L2 instead of 128 + 2.
__device__ __forceinline__ void tc_tile_coords(int t, int tiles_m, int tiles_n, int& tm, int& tn) {
    constexpr int G = 16;
    const int group = G * tiles_n, gid = t / group, first = gid * G;
    const int gsz = min(tiles_m - first, G), r = t - gid * group;
    tm = first + r % gsz;
    tn = r / gsz;
}

struct SgemmTcParams {
    int M, N, kblocks, tiles_m, tiles_n, num_tiles;
    float* C; long long ldc;
    float alpha, beta;
};

__global__ void __launch_bounds__(tcfg::THREADS, 1)
sgemm_tcgen05_3xtf32_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                            const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                            const SgemmTcParams p) {
    using namespace tcfg;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;       // full[s], empty[s], tmem_full[2], tmem_empty[2], tmem ptr
    const uint32_t full0 = bar_base, empty0 = bar_base + 8 * STAGES, tfull0 = bar_base + 16 * STAGES, tempty0 = tfull0 + 16;
    const uint32_t tmem_slot = tempty0 + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full0 + 8 * s, 1); ptx::mbar_init(empty0 + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull0 + 8 * a, 1); ptx::mbar_init(tempty0 + 8 * a, 4); }
        ptx::fence_mbar_init();
    }
    if (warp == 1) {      // TMEM allocation: one warp, address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            ptx::prefetch_tensormap(&map_ahi); ptx::prefetch_tensormap(&map_alo);
            ptx::prefetch_tensormap(&map_bhi); ptx::prefetch_tensormap(&map_blo);
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                int tm, tn;
                tc_tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t full = full0 + 8 * stage, sa = smem_base + stage * STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(full, STAGE_BYTES);
                    ptx::tma_load_2d(sa, &map_ahi, full, kb * BK, tm * BM);
                    ptx::tma_load_2d(sa + TILE_BYTES, &map_alo, full, kb * BK, tm * BM);
                    ptx::tma_load_2d(sa + 2 * TILE_BYTES, &map_bhi, full, kb * BK, tn * BN);
                    ptx::tma_load_2d(sa + 3 * TILE_BYTES, &map_blo, full, kb * BK, tn * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                for (int kb0 = 0; kb0 < p.kblocks; kb0 += KC) {
                    ptx::mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);      // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
                    const int kb1 = min(p.kblocks, kb0 + KC);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        ptx::mbar_wait(full0 + 8 * stage, phase);          // TMA has landed the stage
                        tc_fence_after();
                        const uint32_t sa = smem_base + stage * STAGE_BYTES;
                        const uint64_t d_ahi = tc_desc_kmajor_sw128(sa), d_alo = tc_desc_kmajor_sw128(sa + TILE_BYTES);
                        const uint64_t d_bhi = tc_desc_kmajor_sw128(sa + 2 * TILE_BYTES), d_blo = tc_desc_kmajor_sw128(sa + 3 * TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);    // 32 B per k-step inside the 128 B swizzle row
                            // small terms first, the dominant product last
                            tc_mma_tf32(tmem_d, d_alo + adv, d_bhi + adv, kIdescTf32, ((kb - kb0) | k) != 0);
                            tc_mma_tf32(tmem_d, d_ahi + adv, d_blo + adv, kIdescTf32, 1);
                            tc_mma_tf32(tmem_d, d_ahi + adv, d_bhi + adv, kIdescTf32, 1);
                        }
                        tc_commit(empty0 + 8 * stage);                     // stage free once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                    tc_commit(tfull0 + 8 * acc);                           // chunk complete -> epilogue
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue warps: TMEM -> registers -> C =================
        const int lg = warp & 3;                                           // TMEM lane group this warp may read
        int acc = 0; uint32_t acc_phase = 0;
        const bool use_beta = p.beta != 0.f;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            int tm, tn;
            tc_tile_coords(t, p.tiles_m, p.tiles_n, tm, tn);
            const long long row = (long long)tm * BM + lg * 32 + lane;
            float sum[BN];
            for (int kb0 = 0; kb0 < p.kblocks; kb0 += KC) {
                ptx::mbar_wait(tfull0 + 8 * acc, acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * ACC_COLS + ((uint32_t)(lg * 32) << 16);
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(taddr + c0)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c0 + j] = kb0 == 0 ? __uint_as_float(v[j]) : sum[c0 + j] + __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty0 + 8 * acc);        // the MMA warp may overwrite this accumulator
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (row < p.M) {
#pragma unroll
                for (int j = 0; j < BN; ++j) {
                    const long long col = (long long)tn * BN + j;
                    if (col < p.N) {
                        float* cp = p.C + row + col * p.ldc;
                        float r = p.alpha * sum[j];
                        if (use_beta) r += p.beta * *cp;                   // C is never read when beta == 0
                        *cp = r;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
}

// ---- pack: arbitrary strides -> K-major hi / lo --------------------------------------------------------------------------
// out_hi / out_lo [rows][kpad] (kpad % 32 == 0): element (r, kk) of the logical rows x k operand read at
// src[r * s_row + kk * s_k]; the padding columns are zero.
__global__ void __launch_bounds__(256) sgemm_pack_split_kernel(float* __restrict__ out_hi, float* __restrict__ out_lo, long long kpad,
                                                               const float* __restrict__ src, long long s_row, long long s_k,
                                                               long long rows, long long k) {
    __shared__ float tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32, k0 = (long long)blockIdx.y * 32;
    const bool k_fast = (s_k < 0 ? -s_k : s_k) <= (s_row < 0 ? -s_row : s_row);
    // read coalesced along the source's small stride, write coalesced along k
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = k_fast ? r0 + i : r0 + threadIdx.x, kk = k_fast ? k0 + threadIdx.x : k0 + i;
        float v = 0.f;
        if (r < rows && kk < k) v = src[r * s_row + kk * s_k];
        if (k_fast) tile[i][threadIdx.x] = v; else tile[threadIdx.x][i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i, kk = k0 + threadIdx.x;
        if (r < rows && kk < kpad) {
            const float x = tile[i][threadIdx.x];
            const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            out_hi[r * kpad + kk] = hi;
            out_lo[r * kpad + kk] = x - hi;                              // exact
        }
    }
}

static int make_map_f32_kmajor(CUtensorMap* map, const float* base, size_t rows, size_t kpad) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<Fn>(ptr);
        else (void)cudaGetLastError();
    });
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return NA_ECUDA; }
    cuuint64_t dims[2] = {kpad, rows};
    cuuint64_t strides[1] = {kpad * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)tcfg::BK, (cuuint32_t)tcfg::BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (f32) failed (%d)", (int)r); return NA_ECUDA; }
    return NA_OK;
}

// C (m x n, column-major, ldc) <- alpha * A * B + beta * C;  A logical m x k at (rsa, csa), B logical k x n at (rsb, csb).
int sgemm_tc_colmajor(cudaStream_t s, size_t m, size_t k, size_t n, float alpha, const float* a, ptrdiff_t rsa, ptrdiff_t csa,
                      const float* b, ptrdiff_t rsb, ptrdiff_t csb, float beta, float* c, size_t ldc) {
    using namespace tcfg;
    const size_t kpad = round_up(k, (size_t)BK);
    Scratch ahi, alo, bhi, blo;
    NAB_TRY(ahi.alloc(m * kpad * sizeof(float), s)); NAB_TRY(alo.alloc(m * kpad * sizeof(float), s));
    NAB_TRY(bhi.alloc(n * kpad * sizeof(float), s)); NAB_TRY(blo.alloc(n * kpad * sizeof(float), s));
    {
        dim3 blk(32, 8);
        dim3 ga((unsigned)ceil_div(m, 32), (unsigned)(kpad / 32)), gb((unsigned)ceil_div(n, 32), (unsigned)(kpad / 32));
        if (ga.y > 65535 || gb.y > 65535) { set_error("sgemm: k too large"); return NA_EINVAL; }
        sgemm_pack_split_kernel<<<ga, blk, 0, s>>>(ahi.as<float>(), alo.as<float>(), (long long)kpad, a, rsa, csa, (long long)m, (long long)k);
        NAB_LAUNCH_CHECK();
        sgemm_pack_split_kernel<<<gb, blk, 0, s>>>(bhi.as<float>(), blo.as<float>(), (long long)kpad, b, csb, rsb, (long long)n, (long long)k);
        NAB_LAUNCH_CHECK();
    }
    CUtensorMap mah, mal, mbh, mbl;
    NAB_TRY(make_map_f32_kmajor(&mah, ahi.as<float>(), m, kpad)); NAB_TRY(make_map_f32_kmajor(&mal, alo.as<float>(), m, kpad));
    NAB_TRY(make_map_f32_kmajor(&mbh, bhi.as<float>(), n, kpad)); NAB_TRY(make_map_f32_kmajor(&mbl, blo.as<float>(), n, kpad));
    SgemmTcParams p;
    p.M = (int)m; p.N = (int)n; p.kblocks = (int)(kpad / BK);
    p.tiles_m = (int)ceil_div(m, (size_t)BM); p.tiles_n = (int)ceil_div(n, (size_t)BN);
    p.num_tiles = p.tiles_m * p.tiles_n;
    p.C = c; p.ldc = (long long)ldc; p.alpha = alpha; p.beta = beta;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(sgemm_tcgen05_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); });
    if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(sgemm_tc)", __FILE__, __LINE__);
    const int grid = std::min(p.num_tiles, ctx().sm_count);
    sgemm_tcgen05_3xtf32_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(mah, mal, mbh, mbl, p);
    NAB_LAUNCH_CHECK();
    return NA_OK;
}

}  // namespace nab
