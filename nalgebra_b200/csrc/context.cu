// context.cu -- process context, error capture, memory helpers of libnalgebra_b200.
#include "common.cuh"

namespace nab {

std::atomic<uint64_t> g_launches{0};

char* tls_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error_buffer(), 512, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    set_error("CUDA error %d (%s) in `%s` at %s:%d", (int)e, cudaGetErrorString(e), what, base ? base + 1 : file, line);
    (void)cudaGetLastError();  // clear the sticky-less error so the next call starts clean
    return e == cudaErrorMemoryAllocation ? NA_ENOMEM : NA_ECUDA;
}

Context& ctx() {
    static Context c;
    return c;
}

std::mutex& host_api_mutex() {
    static std::mutex m;
    return m;
}

static std::mutex& init_mutex() {
    static std::mutex m;
    return m;
}

static int init_device(int device) {
    std::lock_guard<std::mutex> lock(init_mutex());
    Context& c = ctx();
    if (c.ready && c.device == device) return NA_OK;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s): libnalgebra_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return NA_ECUDA;
    }
    if (device < 0 || device >= count) {
        set_error("na_init: device %d out of range (have %d)", device, count);
        return NA_EINVAL;
    }
    NAB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NAB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libnalgebra_b200 carries sm_100a code only", device, prop.major, prop.minor);
        return NA_ECUDA;
    }
    if (c.ready) {  // re-bind to another device: drop the old streams
        cudaStreamDestroy(c.stream); cudaStreamDestroy(c.stream2);
        cudaEventDestroy(c.ev_a); cudaEventDestroy(c.ev_b);
        c.ready = false;
    }
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    c.cc_major = prop.major; c.cc_minor = prop.minor;
    NAB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    NAB_CUDA(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
    NAB_CUDA(cudaEventCreateWithFlags(&c.ev_a, cudaEventDisableTiming));
    NAB_CUDA(cudaEventCreateWithFlags(&c.ev_b, cudaEventDisableTiming));
    // keep freed scratch in the pool: the blocked factorizations allocate per call
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    c.ready = true;
    return NA_OK;
}

int ensure_init() {
    Context& c = ctx();
    if (c.ready) {
        // host threads other than the initialising one need the device bound too
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != c.device) NAB_CUDA(cudaSetDevice(c.device));
        return NA_OK;
    }
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cur = 0; (void)cudaGetLastError(); }
    return init_device(cur);
}

int Scratch::alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    if (bytes == 0) bytes = 16;
    NAB_CUDA(cudaMallocAsync(&p, bytes, stream));
    return NA_OK;
}
Scratch::~Scratch() {
    if (p) cudaFreeAsync(p, s);
}

}  // namespace nab

using namespace nab;

extern "C" {

int na_init(int device) { return init_device(device); }

int na_shutdown(void) {
    std::lock_guard<std::mutex> lock(host_api_mutex());
    Context& c = ctx();
    if (!c.ready) return NA_OK;
    cudaStreamSynchronize(c.stream); cudaStreamSynchronize(c.stream2);
    cudaStreamDestroy(c.stream); cudaStreamDestroy(c.stream2);
    cudaEventDestroy(c.ev_a); cudaEventDestroy(c.ev_b);
    c.ready = false;
    return NA_OK;
}

const char* na_last_error(void) { return tls_error_buffer(); }
const char* na_version(void) { return "nalgebra_b200 0.1.0 sm_100a"; }
uint64_t na_kernel_launches(void) { return g_launches.load(); }

int na_dev_malloc(void** ptr, size_t bytes) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaMalloc(ptr, bytes ? bytes : 16));
    return NA_OK;
}
int na_dev_free(void* ptr) {
    NAB_CUDA(cudaFree(ptr));
    return NA_OK;
}
int na_host_alloc_pinned(void** ptr, size_t bytes) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return NA_OK;
}
int na_host_free_pinned(void* ptr) {
    NAB_CUDA(cudaFreeHost(ptr));
    return NA_OK;
}
int na_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return NA_OK;
}
int na_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return NA_OK;
}
// ---- peer memory (multi-GPU GEMM: panels staged over NVLink by the copy engines) ------------------------------
int na_ipc_get_handle(const void* dev_ptr, unsigned char handle[64]) {
    NAB_TRY(ensure_init());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    NAB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle, &h, 64);
    return NA_OK;
}
int na_ipc_open_handle(const unsigned char handle[64], void** dev_ptr) {
    NAB_TRY(ensure_init());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    NAB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return NA_OK;
}
int na_ipc_close_handle(void* dev_ptr) {
    NAB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return NA_OK;
}
int na_memcpy_peer_async(void* dst, const void* src, size_t bytes, void* stream) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return NA_OK;
}

int na_dev_synchronize(void) {
    NAB_TRY(ensure_init());
    NAB_CUDA(cudaDeviceSynchronize());
    return NA_OK;
}

}  // extern "C"
