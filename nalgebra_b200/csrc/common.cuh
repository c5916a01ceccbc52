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
#include <cstdlib>
#include <mutex>
#include <vector>

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

// RAII over the side streams / events of the blocked drivers: an early return (NAB_TRY / NAB_CUDA) between creation
// and the end of a driver neither leaks them nor leaves work pending on them.  Declare guards BEFORE any Scratch
// that allocates on their stream (destruction runs in reverse order).
struct StreamGuard {
    cudaStream_t s = nullptr;
    // high_priority: the greatest priority the device offers (the latency-bound panel chain goes first when CTAs
    // of several streams are pending)
    int create(bool high_priority = false) {
        int lo = 0, hi = 0;
        if (high_priority) NAB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        NAB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : 0));
        return NA_OK;
    }
    ~StreamGuard() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
    operator cudaStream_t() const { return s; }
};
struct EventGuard {
    cudaEvent_t e = nullptr;
    int create() { NAB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return NA_OK; }
    ~EventGuard() { if (e) cudaEventDestroy(e); }
    operator cudaEvent_t() const { return e; }
};

// ---- optional timelines ----------------------------------------------------------------------------
// NAB_LU_TRACE=1 / NAB_CHOL_TRACE=1: start and duration of the phases of a blocked driver, taken with CUDA
// events on the streams the work runs on, printed to stderr when the factorization is done.
struct Timeline {
    bool on = false;
    const char* tag = "";
    cudaEvent_t t0 = nullptr;
    struct Rec { const char* what; size_t j; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> events;
    Timeline(const char* env, const char* tag_) : tag(tag_) { const char* e = getenv(env); on = e && atoi(e) != 0; }
    cudaEvent_t mark(cudaStream_t s) {
        cudaEvent_t e = nullptr;
        if (on) { cudaEventCreate(&e); cudaEventRecord(e, s); events.push_back(e); }
        return e;
    }
    void start(cudaStream_t s) { t0 = mark(s); }
    void add(const char* what, size_t j, cudaEvent_t a, cudaEvent_t b) { if (on) recs.push_back({what, j, a, b}); }
    void dump() {
        if (!on) return;
        cudaDeviceSynchronize();
        for (auto& r : recs) {
            float s0 = 0, d = 0;
            cudaEventElapsedTime(&s0, t0, r.a); cudaEventElapsedTime(&d, r.a, r.b);
            fprintf(stderr, "%s j=%6zu %-7s start %9.3f ms  dur %8.3f ms\n", tag, r.j, r.what, s0, d);
        }
        recs.clear();
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
    }
};

inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return ceil_div(a, b) * b; }

}  // namespace nab
