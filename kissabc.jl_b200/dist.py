"""Multi-rank plumbing (one process per GPU).  torch.distributed is used ONLY to move the 128-byte NCCL id
from rank 0 to the other ranks and for barriers; every data-path collective runs inside libkissabc_cuda.so.

The partition is the one the C ABI uses (kabc_smc_create): contiguous blocks of the GLOBAL particle index,
N/world per rank; Philox counters are keyed by the global index, so results do not depend on `world`.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

NCCL_ID_BYTES = 128


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the particles rank `rank` proposes/simulates/accepts.  n must be a multiple of world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n % world:
        raise ValueError("nparticles must be a multiple of the number of ranks")
    per = n // world
    return per * rank, per * (rank + 1)


def broadcast_id(make_id: Callable[[], bytes], rank: int, device=None) -> bytes:
    """Rank 0 creates the NCCL unique id (make_id), everybody receives it through torch.distributed."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(NCCL_ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        if len(raw) != NCCL_ID_BYTES:
            raise ValueError("NCCL id must be 128 bytes")
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def gloo_exchange(group=None) -> Callable[[bytes], list]:
    """all_gather(bytes) -> list[bytes] over torch.distributed (any backend that moves CPU objects)."""
    import torch.distributed as dist

    def ex(raw: bytes):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, raw, group=group)
        return out
    return ex


def make_context(seed: int, device_index: Optional[int] = None, arena_bytes: int = 0):
    """Context for this process: single-GPU when WORLD_SIZE is 1, else a rank of the job.
    torch.distributed must already be initialised when WORLD_SIZE > 1.  With the nccl backend the 128-byte NCCL id is
    broadcast and the context exchanges its peer arena by itself; with any other backend (gloo: also several ranks on ONE
    GPU) the arena handles travel through torch.distributed and `arena_bytes` must cover the largest handle
    (kabc_smc_arena_bytes / kabc_ais_arena_bytes)."""
    from .api import Context
    rank, world, local = env_rank_world()
    dev = local if device_index is None else device_index
    if world == 1:
        return Context(device=dev, seed=seed)
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        nid = broadcast_id(Context.nccl_unique_id, rank, device=torch.device("cuda", dev))
        return Context(device=dev, seed=seed, rank=rank, world=world, nccl_id=nid)
    return Context(device=dev, seed=seed, rank=rank, world=world, exchange=gloo_exchange(), arena_bytes=arena_bytes)
