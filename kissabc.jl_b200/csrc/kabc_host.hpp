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
};

namespace kabc {

// descriptor ingestion (validates and derives the constants the kernels need)
int ingest_priors(const kabc_prior_t *prior, int d, DPriors &out);
int ingest_model(const kabc_model_t *model, int d, DModel &out);

// host <-> device helpers bound to a context
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
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
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

} // namespace kabc
