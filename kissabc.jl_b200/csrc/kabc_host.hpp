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

constexpr int KABC_MAX_PEERS = 16;
constexpr size_t KABC_ARENA_HEADER = 4096; // barrier flags live at the start of every rank's arena

struct kabc_ctx {
    int device = 0;
    int sm_count = 0;
    uint64_t seed = 0;
    kabc::RoundKeys rk;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr; // only used to move the cudaIpc handles of the arena (and as a host-side barrier)
    long long launches = 0;
    // ---- peer arena: ONE device allocation per rank, mapped into every other rank of the job with cudaIpc (NVLink
    // peer memory) and kept for the life of the context.  smc / ais handles carve their peer-visible buffers out of it
    // at identical offsets on every rank (they are created by identical call sequences).  The first 4 KiB hold the
    // flags of the cross-rank barrier (kabc_peer.cuh); `xseq` counts the barriers this rank has completed.
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0, arena_top = KABC_ARENA_HEADER;
    int arena_users = 0;
    bool arena_attached = false; // peers are mapped (always true when world == 1)
    bool host_exchange = false;  // the host moves the handles (kabc_ctx_arena_export/attach) instead of NCCL
    void *arena_map[KABC_MAX_PEERS] = {};
    unsigned long long *xseq = nullptr;
    // Device-buffer cache: smc/ais handles are created and destroyed once per user call (smc(...), sample(...)); cudaMalloc
    // and above all cudaFree of a few hundred MB cost far more than an smc run, so freed buffers are kept for the next handle.
    // Best fit with bounded slack; idle bytes are capped so that varying sizes cannot pile up.
    struct CacheEntry { void *p; size_t bytes; bool in_use; };
    std::vector<CacheEntry> cache;
    static constexpr size_t IDLE_CAP = (size_t)4 << 30;
    cudaError_t acquire(void **out, size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        CacheEntry *best = nullptr;
        for (auto &e : cache)
            if (!e.in_use && e.bytes >= bytes && e.bytes <= bytes + bytes / 4 + 4096 && (!best || e.bytes < best->bytes)) best = &e;
        if (best) { best->in_use = true; *out = best->p; return cudaSuccess; }
        cudaError_t rc = cudaMalloc(out, bytes);
        if (rc != cudaSuccess) { // out of memory: drop every idle buffer and retry
            cudaGetLastError();
            trim(0);
            rc = cudaMalloc(out, bytes);
        }
        if (rc == cudaSuccess) cache.push_back({*out, bytes, true});
        return rc;
    }
    void release(void *p) {
        for (auto &e : cache)
            if (e.p == p) { e.in_use = false; trim(IDLE_CAP); return; }
        cudaFree(p);
    }
    // free idle buffers (largest first) until at most `keep_idle` idle bytes remain
    void trim(size_t keep_idle) {
        for (;;) {
            size_t idle = 0;
            int big = -1;
            for (size_t q = 0; q < cache.size(); ++q)
                if (!cache[q].in_use) { idle += cache[q].bytes; if (big < 0 || cache[q].bytes > cache[big].bytes) big = (int)q; }
            if (big < 0 || idle <= keep_idle) return;
            cudaFree(cache[big].p);
            cache.erase(cache.begin() + big);
        }
    }
    size_t in_use_count() const {
        size_t n = 0;
        for (auto &e : cache) n += e.in_use ? 1 : 0;
        return n;
    }
};

namespace kabc {

// descriptor ingestion (validates and derives the constants the kernels need)
int ingest_priors(const kabc_prior_t *prior, int d, DPriors &out);
int ingest_model(const kabc_model_t *model, int d, DModel &out);
// costs of the particles named by a device-side list (count in device memory, at most max_count), written to out[i]
// peer arena (kabc_core.cu): make sure the arena holds `bytes` beyond its header and that every peer is mapped; hand out
// `bytes` at an offset that is identical on every rank.  Collective when world > 1.
int arena_reserve(kabc_ctx *ctx, size_t bytes);
int arena_alloc(kabc_ctx *ctx, size_t bytes, size_t *offset);
void arena_release(kabc_ctx *ctx); // one user less; the bump pointer rewinds when nobody is left
// NVTX ranges around the phases of an iteration (header-only NVTX3: no-ops unless a profiler is attached).
// Under a CUDA graph replay the ranges of the captured launches are those of the capture.
void nvtx_push(const char *name);
void nvtx_pop();
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtx_push(name); }
    ~NvtxRange() { nvtx_pop(); }
};
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
