// Shared host/device helpers for the sequitr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sequitr_b200.h"

// ---------------------------------------------------------------- error plumbing
void sq_set_error(const char *fmt, ...);

#define SQ_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            sq_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                         cudaGetErrorString(_e));                                      \
            return SQ_ECUDA;                                                           \
        }                                                                              \
    } while (0)

#define SQ_CHECK_LAUNCH()  SQ_CUDA(cudaGetLastError())

#define SQ_REQUIRE(cond, code, ...)                                                    \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            sq_set_error(__VA_ARGS__);                                                 \
            return (code);                                                             \
        }                                                                              \
    } while (0)

#define SQ_TRY(expr)                                                                   \
    do {                                                                               \
        int _s = (expr);                                                               \
        if (_s != SQ_OK) return _s;                                                    \
    } while (0)

// ---------------------------------------------------------------- handle
struct sq_handle_s {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    // pinned / device staging for the *_host convenience entry points
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    void *dev_arena = nullptr;
    size_t dev_arena_bytes = 0;
    cudaStream_t stream = nullptr;      // library-owned stream for *_host calls
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // the *_host entry points share the staging arena, the two streams and the events above: they
    // serialise on this mutex (device-pointer entry points own nothing in the handle and need no lock)
    std::mutex host_mu;
};

// Scope of one *_host call: serialises the calls of a handle and, on EVERY exit path (also the early
// error returns), drains both library streams so that no asynchronous copy still reads or writes the
// caller's host buffers after the call has returned.
struct SqHostCall {
    sq_handle_s *h;
    std::lock_guard<std::mutex> lock;
    explicit SqHostCall(sq_handle_s *hh) : h(hh), lock(hh->host_mu) {}
    ~SqHostCall()
    {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
    }
};

int sq_reserve_pinned(sq_handle_s *h, size_t bytes);
int sq_reserve_device(sq_handle_s *h, size_t bytes);

static inline size_t sq_align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct SqArena {
    char *base;
    size_t size, off;
    SqArena(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <typename T> T *take(size_t count) {
        size_t bytes = sq_align_up(count * sizeof(T));
        T *r = (T *)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

static inline int sq_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
