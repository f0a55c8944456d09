// kabc_host.hpp -- host-side plumbing shared by the translation units of libkissabc_cuda.so
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "kabc_device.cuh"
#include "kabc_models.cuh"

typedef struct ncclComm *ncclComm_t;

namespace kabc {

extern thread_local std::string g_last_error;
int set_error(int code, const char *fmt, ...);

#define KABC_CUDA_TRY(expr)                                                                                  \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return ::kabc::set_error(KABC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                                     __FILE__, __LINE__);                                                    \
    } while (0)

} // namespace kabc

struct kabc_ctx {
    int device = 0;
    int sm_count = 0;
    uint64_t seed = 0;
    kabc::RoundKeys rk;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    long long launches = 0;
    // Device-buffer cache: smc/ais handles are created and destroyed once per user call (smc(...), sample(...)); cudaMalloc
    // and above all cudaFree of a few hundred MB cost far more than an smc run, so freed buffers are kept for the next handle.
    struct CacheEntry { void *p; size_t bytes; bool in_use; };
    std::vector<CacheEntry> cache;
    cudaError_t acquire(void **out, size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (auto &e : cache)
            if (!e.in_use && e.bytes == bytes) { e.in_use = true; *out = e.p; return cudaSuccess; }
        cudaError_t rc = cudaMalloc(out, bytes);
        if (rc != cudaSuccess) { // out of memory: drop every idle buffer and retry
            cudaGetLastError();
            trim();
            rc = cudaMalloc(out, bytes);
        }
        if (rc == cudaSuccess) cache.push_back({*out, bytes, true});
        return rc;
    }
    void release(void *p) {
        for (auto &e : cache)
            if (e.p == p) { e.in_use = false; return; }
        cudaFree(p);
    }
    void trim() {
        std::vector<CacheEntry> keep;
        for (auto &e : cache) {
            if (e.in_use) keep.push_back(e);
            else cudaFree(e.p);
        }
        cache.swap(keep);
    }
};

namespace kabc {

// descriptor ingestion (validates and derives the constants the kernels need)
int ingest_priors(const kabc_prior_t *prior, int d, DPriors &out);
int ingest_model(const kabc_model_t *model, int d, DModel &out);
// costs of the particles named by a device-side list (count in device memory, at most max_count), written to out[i]
int eval_cost_list_device(kabc_ctx *ctx, const DModel &m, const double *d_th, long long N, const unsigned int *list,
                          const unsigned int *count, long long max_count, uint32_t tag, uint32_t epoch, double *d_out);

// host <-> device helpers bound to a context
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    kabc_ctx *owner = nullptr; // non-null: the buffer comes from (and returns to) the context's cache
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    cudaError_t alloc(kabc_ctx *ctx, size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        cudaError_t rc = ctx->acquire((void **)&p, count * sizeof(T));
        if (rc == cudaSuccess) owner = ctx;
        return rc;
    }
    void release() {
        if (p) {
            if (owner) owner->release(p);
            else cudaFree(p);
        }
        p = nullptr;
        n = 0;
        owner = nullptr;
    }
};

} // namespace kabc
