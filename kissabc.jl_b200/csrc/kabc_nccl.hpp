// kabc_nccl.hpp -- NCCL is bound lazily with dlopen("libnccl.so.2") the first time a multi-rank context is
// created, so that (a) single-GPU users need no NCCL at all and (b) inside a PyTorch host process the
// already-loaded torch-bundled NCCL is reused instead of a second copy.
#pragma once
#include <cstddef>
#include <cstdint>
#include "kabc_host.hpp"

namespace kabc {
int nccl_unique_id(char id[KABC_NCCL_ID_BYTES]);
int nccl_comm_init(kabc_ctx *ctx, const char id[KABC_NCCL_ID_BYTES]);
void nccl_comm_destroy(kabc_ctx *ctx);
// in-place all-gather: every rank owns `count_per_rank` elements at offset rank*count_per_rank of `buf`
int nccl_allgather_inplace(kabc_ctx *ctx, void *buf, size_t bytes_per_rank);
int nccl_allreduce_sum_u64(kabc_ctx *ctx, unsigned long long *buf, size_t count);
int nccl_group_start();
int nccl_group_end();
} // namespace kabc
