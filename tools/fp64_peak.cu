// FP64 peak micro-benchmark for B200 (sm_100a): DFMA chains vs DMMA (mma.sync f64) chains.
// Step 0 of the build plan: decides the DGEMM inner loop and gives the roofline
// denominator `fp64_tflops`.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// NACC independent accumulator tiles per warp.
template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-12;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC, int KSHAPE>
__global__ void __launch_bounds__(256) dmma16_kernel(double* out, int iters) {
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i * 1e-10;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-12 + i * 1e-11;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 2 * i; c[i][3] = 3 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (KSHAPE == 4) dmma1684(c[i], a, b[0]);
            else if (KSHAPE == 8) dmma1688(c[i], a, b);
            else dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// DMMA m8n8k4 fed from shared memory (LDS.64 per fragment), to see whether LDS traffic co-issues.
template <int WM8, int WN8>   // warp tile = (8*WM8) x (8*WN8)
__global__ void __launch_bounds__(256) dmma884_lds_kernel(double* out, int iters) {
    __shared__ double sa[16 * 132], sb[16 * 132];
    for (int i = threadIdx.x; i < 16 * 132; i += blockDim.x) { sa[i] = i * 1e-9; sb[i] = 1.0 + i * 1e-12; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double c[WM8][WN8][2];
#pragma unroll
    for (int i = 0; i < WM8; ++i)
#pragma unroll
        for (int j = 0; j < WN8; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // 4 k-steps of 4 out of a 16-deep stage
            double a[WM8], b[WN8];
#pragma unroll
            for (int i = 0; i < WM8; ++i) a[i] = sa[(kk * 4 + (lane & 3)) * 132 + ((warp & 1) * 64 + i * 8 + (lane >> 2)) % 128];
#pragma unroll
            for (int j = 0; j < WN8; ++j) b[j] = sb[(kk * 4 + (lane & 3)) * 132 + ((warp >> 1) * 32 + j * 8 + (lane >> 2)) % 128];
#pragma unroll
            for (int i = 0; i < WM8; ++i)
#pragma unroll
                for (int j = 0; j < WN8; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < WM8; ++i)
#pragma unroll
        for (int j = 0; j < WN8; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000, reps = 20;
    for (int cps = 1; cps <= 4; cps *= 2) {   // CTAs (256 thr) per SM
        int grid = sms * cps;
        double threads = (double)grid * 256, warps = threads / 32;
#define REPORT(name, flops, ms) printf("{\"kernel\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", name, cps, ms, (flops) / (ms * 1e-3) / 1e12); fflush(stdout)
        { double ms = time_ms([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
          REPORT("dfma_acc8", threads * 8.0 * iters * 2, ms); }
        { double ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, reps);
          REPORT("dfma_acc16", threads * 16.0 * iters * 2, ms); }
        { double ms = time_ms([&] { dmma884_kernel<4><<<grid, 256>>>(out, iters); }, reps);
          REPORT("dmma_m8n8k4_acc4", warps * 4.0 * iters * 2 * 256, ms); }
        { double ms = time_ms([&] { dmma884_kernel<16><<<grid, 256>>>(out, iters); }, reps);
          REPORT("dmma_m8n8k4_acc16", warps * 16.0 * iters * 2 * 256, ms); }
        { double ms = time_ms([&] { dmma16_kernel<8, 4><<<grid, 256>>>(out, iters); }, reps);
          REPORT("dmma_m16n8k4_acc8", warps * 8.0 * iters * 2 * 512, ms); }
        { double ms = time_ms([&] { dmma16_kernel<8, 8><<<grid, 256>>>(out, iters / 2); }, reps);
          REPORT("dmma_m16n8k8_acc8", warps * 8.0 * (iters / 2) * 2 * 1024, ms); }
        { double ms = time_ms([&] { dmma16_kernel<8, 16><<<grid, 256>>>(out, iters / 4); }, reps);
          REPORT("dmma_m16n8k16_acc8", warps * 8.0 * (iters / 4) * 2 * 2048, ms); }
        if (cps == 1) {
            { double ms = time_ms([&] { dmma884_lds_kernel<8, 4><<<grid, 256>>>(out, iters / 16); }, reps);
              REPORT("dmma_m8n8k4_lds_64x32", warps * 32.0 * 4 * (iters / 16) * 2 * 256, ms); }
            { double ms = time_ms([&] { dmma884_lds_kernel<4, 4><<<grid, 256>>>(out, iters / 16); }, reps);
              REPORT("dmma_m8n8k4_lds_32x32", warps * 16.0 * 4 * (iters / 16) * 2 * 256, ms); }
        }
    }
    // Sustained: ~3 s of the best-known shape to see the power-capped clock.
    {
        int grid = sms * 2;
        double warps = (double)grid * 8;
        double ms = time_ms([&] { dmma884_kernel<16><<<grid, 256>>>(out, iters); }, 400);
        printf("{\"kernel\": \"dmma_m8n8k4_acc16_sustained\", \"ms\": %.4f, \"tflops\": %.3f}\n", ms, warps * 16.0 * iters * 2 * 256 / (ms * 1e-3) / 1e12);
        double threads = (double)grid * 256;
        ms = time_ms([&] { dfma_kernel<16><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 400);
        printf("{\"kernel\": \"dfma_acc16_sustained\", \"ms\": %.4f, \"tflops\": %.3f}\n", ms, threads * 16.0 * iters * 2 / (ms * 1e-3) / 1e12);
    }
    return 0;
}
