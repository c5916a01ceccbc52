// Inter-CTA signalling latency on B200: 2 CTAs ping-pong a sequence number through global memory.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <int MODE> __device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
    if (MODE == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 1) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 2) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (MODE == 3) atomicExch(p, v);
    if (MODE == 4) asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE> __device__ __forceinline__ unsigned long long ld_flag(unsigned long long* p) {
    unsigned long long v;
    if (MODE == 0) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 1) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 2) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (MODE == 3) v = atomicAdd(p, 0ull);
    if (MODE == 4) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// flags[0] written by CTA 0, flags[16] (different 128B line) by CTA 1.
template <int MODE> __global__ void pingpong(unsigned long long* flags, int iters, long long* out) {
    if (threadIdx.x != 0) return;
    const int me = blockIdx.x;
    unsigned long long* mine = flags + me * 16;
    unsigned long long* other = flags + (1 - me) * 16;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (me == 0) { st_flag<MODE>(mine, i); while (ld_flag<MODE>(other) < (unsigned long long)i) {} }
        else { while (ld_flag<MODE>(other) < (unsigned long long)i) {} st_flag<MODE>(mine, i); }
    }
    if (me == 0) out[0] = clock64() - t0;
}
// own-store visibility: st then poll own flag
template <int MODE> __global__ void selfpoll(unsigned long long* flags, int iters, long long* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) { st_flag<MODE>(flags + 32, i); while (ld_flag<MODE>(flags + 32) < (unsigned long long)i) {} }
    out[1] = clock64() - t0;
}
// all-to-all: G CTAs each publish one word, every CTA polls all G words (like the panel exchange)
template <int MODE> __global__ void alltoall(unsigned long long* flags, int iters, long long* out) {
    const int G = gridDim.x, me = blockIdx.x, tid = threadIdx.x;
    long long t0 = clock64();
    for (int i = 1; i <= iters; ++i) {
        if (tid == 0) st_flag<MODE>(flags + 64 + me * 16, i);
        for (int g = tid; g < G; g += blockDim.x) while (ld_flag<MODE>(flags + 64 + g * 16) < (unsigned long long)i) {}
        __syncthreads();
    }
    if (me == 0 && tid == 0) out[2] = clock64() - t0;
}
template <int MODE> int run(const char* name, unsigned long long* flags, long long* out) {
    const int iters = 2000;
    long long h[3];
    CK(cudaMemset(flags, 0, 8 * 4096));
    void* args[] = {&flags, (void*)&iters, &out};
    CK(cudaLaunchCooperativeKernel((void*)pingpong<MODE>, 2, 32, args, 0, 0));
    selfpoll<MODE><<<1, 32>>>(flags, iters, out);
    CK(cudaDeviceSynchronize());
    printf("%-28s ping-pong round trip", name);
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    printf(" %6lld cyc   self st->ld %6lld cyc", h[0] / iters, h[1] / iters);
    for (int G : {8, 32, 86, 148}) {
        CK(cudaMemset(flags, 0, 8 * 4096));
        CK(cudaLaunchCooperativeKernel((void*)alltoall<MODE>, G, 128, args, 0, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("   all2all G=%d: %6lld", G, h[2] / iters);
    }
    printf("\n");
    return 0;
}
int main() {
    unsigned long long* flags; long long* out;
    CK(cudaMalloc(&flags, 8 * 4096)); CK(cudaMalloc(&out, 64));
    run<0>("volatile", flags, out);
    run<1>("relaxed.gpu", flags, out);
    run<2>("release/acquire.gpu", flags, out);
    run<3>("atomicExch/atomicAdd0", flags, out);
    run<4>("st.cg/ld.cg", flags, out);
    return 0;
}
