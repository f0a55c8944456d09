"""N>1 host logic on CPU: world_size-2 `gloo` processes.

(1) the NCCL-id broadcast plumbing and the shard partition of kissabc.jl_b200/dist.py;
(2) a multi-rank smc SCHEDULE (sharded sweep: every rank proposes / simulates / accepts its own particles against the
    pre-sweep ensemble; rows and summed counters exchanged after the sweep; quantile / cut / resample decided from replicated
    scalars) emulated with the CPU oracle: the 2-rank run must reproduce the single-process run bit for bit -- the property
    the device path relies on (kabc_smc.cu shards the state as well and reads rows from their owners; its own multi-rank
    runs are checked on the GPU, tests/test_gpu_multi.py, with ranks sharing one GPU when only one is visible).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kissabc_jl_b200 as k
    from oracle import oracle as O
    from common import SEED

    # (1) id broadcast + partition
    nid = k.dist.broadcast_id(lambda: bytes(range(128)), rank)
    assert nid == bytes(range(128))
    assert k.dist.env_rank_world() == (rank, world, rank)
    N = 600
    lo, hi = k.dist.shard_range(N, rank, world)
    assert hi - lo == N // world

    # (2) sharded smc schedule with the oracle
    pri = O.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)])
    mod = O.make_model(O.NORMAL_MEANSTD, 50, (2.0, 0.04), (50.0,))
    cfg = O.smc_config(nparticles=N, alpha=0.8, min_r_ess=0.5, mcmc_retrys=2, mcmc_tol=0.4, max_iterations=8)
    s = O.Smc(SEED, pri, mod, cfg)
    s.init()  # (the device shards the init too; values depend only on the global id)

    def allgather_rows():
        th, X, lpi, alive = s.state()
        parts = [None] * world
        dist.all_gather_object(parts, (th[:, lo:hi].copy(), X[lo:hi].copy(), lpi[lo:hi].copy()))
        for r, (t, x, l) in enumerate(parts):
            a, b = k.dist.shard_range(N, r, world)
            th[:, a:b], X[a:b], lpi[a:b] = t, x, l
        s.set_state(th, X, lpi, alive)

    stops = []
    for _ in range(8):
        s.cut()
        for r in range(1 + 2):
            acc, ev, events = s.sweep_range(lo, hi)
            allgather_rows()
            t = torch.tensor([acc, ev, events], dtype=torch.int64)
            dist.all_reduce(t)
            if s.sweep_commit(*[int(v) for v in t]):
                break
        stops.append(s.finish())
        if stops[-1]:
            break
    th, X, lpi, alive = s.state()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), th=th, X=X, lpi=lpi, alive=alive, eps=s.scalars()["eps"],
             evals=s.scalars()["cost_evals"], stops=np.array(stops))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_schedule_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    from common import SEED
    O.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    # single-process reference
    pri = O.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)])
    mod = O.make_model(O.NORMAL_MEANSTD, 50, (2.0, 0.04), (50.0,))
    s = O.Smc(SEED, pri, mod, O.smc_config(nparticles=600, alpha=0.8, min_r_ess=0.5, mcmc_retrys=2, mcmc_tol=0.4, max_iterations=8))
    s.init()
    stops = []
    for _ in range(8):
        stops.append(s.iterate())
        if stops[-1]:
            break
    th, X, lpi, alive = s.state()
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert (z["th"].view(np.uint64) == th.view(np.uint64)).all()
        assert (z["X"].view(np.uint64) == X.view(np.uint64)).all()
        assert (z["lpi"].view(np.uint64) == lpi.view(np.uint64)).all()
        assert (z["alive"] == alive).all()
        assert float(z["eps"]) == s.scalars()["eps"] and int(z["evals"]) == s.scalars()["cost_evals"]
        assert list(z["stops"]) == stops


def test_shard_range_partition():
    import kissabc_jl_b200 as k
    for n, w in [(8, 1), (8, 2), (1 << 20, 8), (24, 4)]:
        ranges = [k.dist.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    with pytest.raises(ValueError):
        k.dist.shard_range(10, 0, 4)
    with pytest.raises(ValueError):
        k.dist.shard_range(8, 2, 2)
