// Runtime entry points of the C ABI: handles, errors, staging arenas.
#include "sq_common.cuh"

static thread_local char g_err[1024] = "";

void sq_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *sq_version(void) { return "sequitr_b200 0.1 (sm_100a)"; }

extern "C" const char *sq_last_error(void) { return g_err; }

extern "C" int sq_destroy(sq_handle_t h);

extern "C" int sq_create(int device, sq_handle_t *out)
{
    SQ_REQUIRE(out, SQ_EINVAL, "sq_create: null pointer");
    int count = 0;
    SQ_CUDA(cudaGetDeviceCount(&count));
    SQ_REQUIRE(device >= 0 && device < count, SQ_EINVAL, "sq_create: device %d of %d", device, count);
    SQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SQ_CUDA(cudaGetDeviceProperties(&prop, device));
    // The library is compiled for sm_100a only: fail loudly anywhere else.
    SQ_REQUIRE(prop.major == 10, SQ_EUNSUPPORTED,
               "sq_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
               device, prop.major, prop.minor);
    sq_handle_s *h = new sq_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    h->total_mem = prop.totalGlobalMem;
    // a failure below must not leak the half-built handle (sq_destroy frees whatever exists)
    auto build = [&]() -> int {
        SQ_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        SQ_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            SQ_CUDA(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
            SQ_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
        return SQ_OK;
    };
    const int st = build();
    if (st != SQ_OK) { sq_destroy(h); return st; }
    *out = h;
    return SQ_OK;
}

extern "C" int sq_destroy(sq_handle_t h)
{
    if (!h) return SQ_OK;
    cudaSetDevice(h->device);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->dev_arena) cudaFree(h->dev_arena);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    }
    delete h;
    return SQ_OK;
}

extern "C" int sq_device_info(sq_handle_t h, int *sm_count, int *cc_major, int *cc_minor,
                              size_t *total_mem)
{
    SQ_REQUIRE(h, SQ_EINVAL, "sq_device_info: null handle");
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    if (total_mem) *total_mem = h->total_mem;
    return SQ_OK;
}

extern "C" int sq_host_register(void *ptr, size_t bytes)
{
    SQ_REQUIRE(ptr && bytes, SQ_EINVAL, "sq_host_register: null buffer");
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();                       // do not leave a sticky error for the caller's next CUDA call
        sq_set_error("sq_host_register: %zu bytes -> %s", bytes, cudaGetErrorString(e));
        return SQ_ECUDA;
    }
    return SQ_OK;
}

extern "C" int sq_host_unregister(void *ptr)
{
    SQ_REQUIRE(ptr, SQ_EINVAL, "sq_host_unregister: null buffer");
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        sq_set_error("sq_host_unregister: %s", cudaGetErrorString(e));
        return SQ_ECUDA;
    }
    return SQ_OK;
}

int sq_reserve_pinned(sq_handle_s *h, size_t bytes)
{
    if (bytes <= h->pinned_bytes) return SQ_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
    SQ_CUDA(cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return SQ_OK;
}

int sq_reserve_device(sq_handle_s *h, size_t bytes)
{
    if (bytes <= h->dev_arena_bytes) return SQ_OK;
    if (h->dev_arena) {
        SQ_CUDA(cudaDeviceSynchronize());
        cudaFree(h->dev_arena);
    }
    h->dev_arena = nullptr;
    h->dev_arena_bytes = 0;
    SQ_CUDA(cudaMalloc(&h->dev_arena, bytes));
    h->dev_arena_bytes = bytes;
    return SQ_OK;
}
