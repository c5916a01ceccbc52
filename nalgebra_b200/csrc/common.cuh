// common.cuh -- shared host-side plumbing of libnalgebra_b200: status codes, error capture,
// the process context (device + internal stream), launch accounting, small helpers.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/nalgebra_b200.h"

namespace nab {

// ---- error capture -----------------------------------------------------------------------------
char* tls_error_buffer();                     // 512 bytes, thread local
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define NAB_CUDA(expr)                                                            \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) return ::nab::cuda_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define NAB_TRY(expr)                       \
    do {                                    \
        int s__ = (expr);                   \
        if (s__ < 0) return s__;            \
    } while (0)

// ---- context -----------------------------------------------------------------------------------
struct Context {
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    cudaStream_t stream = nullptr;       // internal stream used by host-pointer entry points
    cudaStream_t stream2 = nullptr;      // second stream: panel / update overlap
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    bool ready = false;
};
Context& ctx();
int ensure_init();                          // lazy na_init(current device or 0)
std::mutex& host_api_mutex();               // serialises host-pointer entry points

extern std::atomic<uint64_t> g_launches;    // kernels launched by this library
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Checks the launch that was just enqueued.
#define NAB_LAUNCH_CHECK()                                                        \
    do {                                                                          \
        ::nab::count_launch();                                                    \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) return ::nab::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// ---- stream-ordered scratch memory ---------------------------------------------------------------
// RAII over cudaMallocAsync/cudaFreeAsync on one stream.
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    int alloc(size_t bytes, cudaStream_t stream);
    ~Scratch();
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return ceil_div(a, b) * b; }

}  // namespace nab
