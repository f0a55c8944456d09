"""Multi-GPU smc (needs >= 2 GPUs; skipped on a 1-GPU box): a G-rank run must reproduce the 1-GPU run bit for bit
on every rank -- the Philox counters are keyed by the global particle id and all control is replicated."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


# retrys = 1: retry sweeps, launched kernel by kernel; retrys = 0: the iteration is replayed from a CUDA graph that also holds
# the NCCL all-gather.  packed = "1"/"0" forces the packed inbox / direct row pushes of the rows finalised by propose
# (default: packed from 4 ranks up); "" keeps the default.
@pytest.mark.parametrize("name,prec,N,iters,retrys,packed", [
    ("normal_small", "f64", 4096, 8, 1, ""), ("normal_smc", "f32", 1 << 15, 6, 1, ""), ("lv_smc", "f64", 512, 4, 1, ""),
    ("normal_small", "f64", 4096, 8, 0, "1"), ("normal_small", "f64", 4096, 8, 0, "0"), ("normal_small", "f64", 4096, 8, 1, "1"),
])
def test_multi_rank_equals_single_gpu(kabc, ctx, tmp_path, name, prec, N, iters, retrys, packed):
    g = _ngpu()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if g < 4 else 4
    port = 29500 + os.getpid() % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py"), str(tmp_path), name, prec, str(N), str(iters),
           str(retrys)]
    env = dict(os.environ)
    if packed:
        env["KABC_PACKED_PUSH"] = packed
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    prior, cost = kabc.workloads.WORKLOADS[name](prec) if name != "normal_small" else kabc.workloads.normal(prec, 100)
    s = kabc.SmcSession(ctx, prior, cost, kabc.smc_config(nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_retrys=retrys, mcmc_tol=0.3, max_iterations=iters))
    s.init()
    stops = []
    for _ in range(iters):
        stops.append(s.iterate())
        if stops[-1]:
            break
    th, X, lpi, alive = s.state()
    sc = s.scalars()
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert (z["th"].view(np.uint64) == th.view(np.uint64)).all(), f"theta differs on rank {rank}"
        assert (z["X"].view(np.uint64) == X.view(np.uint64)).all() and (z["lpi"].view(np.uint64) == lpi.view(np.uint64)).all()
        assert (z["alive"] == alive).all() and float(z["eps"]) == sc["eps"]
        assert int(z["evals"]) == sc["cost_evals"] and int(z["accepted"]) == sc["accepted"] and int(z["events"]) == sc["events"]
        assert list(z["stops"]) == stops
