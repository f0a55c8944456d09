// kabc_nccl.cu -- lazy NCCL binding (see kabc_nccl.hpp)
#include <dlfcn.h>
#include "kabc_nccl.hpp"

namespace kabc {

// minimal NCCL ABI (stable since 2.x): see /usr/include/nccl.h
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult;
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_SUM = 0 };

static struct NcclApi {
    void *handle = nullptr;
    ncclResult (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    ncclResult (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
    ncclResult (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult (*GroupStart)() = nullptr;
    ncclResult (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult) = nullptr;
} g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return KABC_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(KABC_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define KABC_SYM(field, name)                                                                        \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                                      \
    if (!g_nccl.field) return set_error(KABC_ERR_NCCL, "libnccl lacks symbol %s", name);
    KABC_SYM(GetUniqueId, "ncclGetUniqueId")
    KABC_SYM(CommInitRank, "ncclCommInitRank")
    KABC_SYM(CommDestroy, "ncclCommDestroy")
    KABC_SYM(AllGather, "ncclAllGather")
    KABC_SYM(AllReduce, "ncclAllReduce")
    KABC_SYM(GroupStart, "ncclGroupStart")
    KABC_SYM(GroupEnd, "ncclGroupEnd")
    KABC_SYM(GetErrorString, "ncclGetErrorString")
#undef KABC_SYM
    g_nccl.handle = h;
    return KABC_OK;
}

#define KABC_NCCL_TRY(expr)                                                                          \
    do {                                                                                             \
        ncclResult _r = (expr);                                                                      \
        if (_r != 0) return set_error(KABC_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(_r)); \
    } while (0)

int nccl_unique_id(char id[KABC_NCCL_ID_BYTES]) {
    if (!id) return set_error(KABC_ERR_INVALID_ARG, "id is NULL");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId_t u;
    KABC_NCCL_TRY(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, KABC_NCCL_ID_BYTES);
    return KABC_OK;
}

int nccl_comm_init(kabc_ctx *ctx, const char id[KABC_NCCL_ID_BYTES]) {
    if (!id) return set_error(KABC_ERR_INVALID_ARG, "NCCL unique id is NULL");
    if (int rc = nccl_load()) return rc;
    ncclUniqueId_t u;
    memcpy(u.internal, id, KABC_NCCL_ID_BYTES);
    KABC_NCCL_TRY(g_nccl.CommInitRank(&ctx->comm, ctx->world, u, ctx->rank));
    return KABC_OK;
}

void nccl_comm_destroy(kabc_ctx *ctx) {
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
}

int nccl_allgather_inplace(kabc_ctx *ctx, void *buf, size_t bytes_per_rank) {
    const char *send = (const char *)buf + (size_t)ctx->rank * bytes_per_rank;
    KABC_NCCL_TRY(g_nccl.AllGather(send, buf, bytes_per_rank, NCCL_UINT8, ctx->comm, ctx->stream));
    return KABC_OK;
}

int nccl_allreduce_sum_u64(kabc_ctx *ctx, unsigned long long *buf, size_t count) {
    KABC_NCCL_TRY(g_nccl.AllReduce(buf, buf, count, NCCL_UINT64, NCCL_SUM, ctx->comm, ctx->stream));
    return KABC_OK;
}

int nccl_group_start() { KABC_NCCL_TRY(g_nccl.GroupStart()); return KABC_OK; }
int nccl_group_end() { KABC_NCCL_TRY(g_nccl.GroupEnd()); return KABC_OK; }

} // namespace kabc
