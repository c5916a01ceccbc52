// FP64 vector-pipe instruction rates on B200: DFMA vs separate DMUL + DADD (unfused, what the LU leaf must use).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_rate dp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double* out, double a, double b, int iters) {
    double x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) x[i] = fma(a, b, x[i]);
            if (MODE == 1) x[i] = __dadd_rn(__dmul_rn(a, x[(i + 1) & 15]), x[i]);
            if (MODE == 2) x[i] = __dmul_rn(a, x[i]);
            if (MODE == 3) x[i] = __dadd_rn(b, x[i]);
        }
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int threads, int ops_per_iter) {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(out, 1.0000001, 1e-9, 10);
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(out, 1.0000001, 1e-9, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = 148.0 * threads * iters * 16 * ops_per_iter;
    printf("%-22s threads/SM %4d: %.1f G DP-instr-lanes/s = %.1f lanes/clk/SM (at 1.965 GHz)\n", name, threads, instr / ms / 1e6, instr / ms / 1e6 / 148 / 1.965);
    cudaFree(out);
}
int main() {
    for (int t : {256, 512, 1024}) {
        run<0>("DFMA", t, 1); run<1>("DMUL+DADD (dependent)", t, 2); run<2>("DMUL", t, 1); run<3>("DADD", t, 1);
    }
    return 0;
}
