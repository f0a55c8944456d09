// kabc_peer.cuh -- cross-rank synchronisation through peer memory (one process per GPU, NVLink / NVSwitch).
//
// Every rank owns an ARENA (one device allocation, kabc_ctx::arena) that is mapped into every other rank with cudaIpc.
// All ranks of a job execute the same kernel sequence (the control flow of smc / AIS is decided from replicated
// scalars), so a barrier is just a monotone sequence number:
//     arrive:  fence.sys ; write  seq+1  into slot [my rank] of the flag array at the start of EVERY peer's arena
//     wait:    spin until slot [r] of MY flag array holds >= seq+1 for every r
// Data written before the arrival -- into the own arena (read by the peers afterwards) or pushed into a peer's arena --
// is visible to a rank once it has observed the flag (release / acquire at system scope).  A barrier is executed by
// ONE thread block per rank: the last block of the kernel that produced the data (last_block below), so the consumer
// kernel, which follows in stream order, needs no prologue.  No NCCL call, no host round trip, ~2 NVLink latencies.
//
// A rank that fails (error flag set, peer gone) never arrives: waits are bounded by a timeout after which the waiter
// raises KABC_ERR_PEER in its own control block and every later kernel becomes a no-op.
#pragma once
#include <cuda_runtime.h>
#include "kabc_host.hpp"

namespace kabc {

struct XPeer {
    unsigned char *arena[KABC_MAX_PEERS]; // arena base of rank r as mapped into this process (arena[rank] = the local one)
    unsigned long long *seq;              // barriers this rank has completed (device memory of the context)
    int rank, world;
};

inline XPeer make_xpeer(const kabc_ctx *ctx) {
    XPeer x;
    for (int r = 0; r < KABC_MAX_PEERS; ++r) x.arena[r] = (unsigned char *)ctx->arena_map[r];
    x.seq = ctx->xseq;
    x.rank = ctx->rank;
    x.world = ctx->world;
    return x;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// true for exactly one block of the grid: the last one to arrive (its reads see every other block's writes).
// sys: the other blocks wrote into PEER memory, which the last block is about to publish with a barrier.
// One fence per block, by the thread that takes the ticket AFTER the block-wide barrier (the pattern of a cooperative
// grid sync): fences are cumulative over what the barrier ordered before them, and a fence.sys costs microseconds -- one
// per warp of a 512-thread block was most of a barrier's latency.
__device__ __forceinline__ bool last_block(unsigned int *ticket, bool sys = false) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sys) __threadfence_system();
        else __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    return s_last;
}

#ifndef KABC_PEER_TIMEOUT_NS
#define KABC_PEER_TIMEOUT_NS 20000000000ull // 20 s: ranks sharing one GPU (tests) are time-sliced by the driver
#endif

// Cross-rank barrier, called by EVERY thread of ONE block per rank (blockDim.x >= world).  Returns false on timeout.
// Thread r (r < world, all in warp 0) handles peer r: fence.sys (cumulative over the block's earlier writes, which the
// block barrier ordered before it), release-store of the sequence number into the peer's flag array, acquire-spin on the
// peer's slot of the own flag array; the closing block barrier orders every thread's later reads after the acquires.
__device__ __forceinline__ bool xbarrier(const XPeer &x) {
    if (x.world == 1) { // single rank: still a block-level barrier (callers read what other threads wrote before it)
        __syncthreads();
        return true;
    }
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    const unsigned long long n = *x.seq + 1ull;
    const int r = (int)threadIdx.x;
    if (r < x.world && r != x.rank) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long *>(x.arena[r]) + x.rank, n);
        const unsigned long long *f = reinterpret_cast<const unsigned long long *>(x.arena[x.rank]) + r;
        const unsigned long long t0 = global_timer_ns();
        unsigned int spins = 0;
        while (ld_acquire_sys(f) < n) {
            if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > KABC_PEER_TIMEOUT_NS) { s_ok = 0; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) *x.seq = n; // also after a timeout: the job is dead anyway, keep the counters aligned
    __syncthreads();
    return s_ok != 0;
}

} // namespace kabc
