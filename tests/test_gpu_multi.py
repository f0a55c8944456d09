"""Multi-rank smc / AIS: a G-rank run must reproduce the 1-rank run bit for bit on every rank -- the Philox counters are keyed
by the global particle id, the control flow is decided from replicated scalars, and ranks only exchange rows and counters
through NVLink peer memory (flag barriers, no NCCL on the data path).

Two launch modes.  "gloo": the ranks SHARE the visible GPU(s) (rank % device_count) and the arena handles travel through
torch.distributed -- this runs on a 1-GPU box too (the driver time-slices the ranks), so the multi-rank code path is always
exercised.  "nccl": one GPU per rank (needs >= 2 GPUs): the production path, NVLink between the ranks."""
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


def _launch(tmp_path, world, mode, name, prec, N, iters, retrys, backend, env_extra=None):
    port = 29500 + (os.getpid() * 7 + world) % 1000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_worker.py"), str(tmp_path), mode, name, prec, str(N),
           str(iters), str(retrys), backend]
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _worlds(backend):
    g = _ngpu()
    if backend == "gloo":
        return [2, 3]
    ws = [w for w in (2, 4, 8) if w <= g]
    only = os.environ.get("KABC_TEST_WORLDS")  # e.g. "8": a single world size on an expensive multi-GPU lease
    if only:
        ws = [w for w in ws if str(w) in only.split(",")]
    return ws


def _check_smc(kabc, ctx, tmp_path, world, name, prec, N, iters, retrys):
    prior, cost = kabc.workloads.WORKLOADS[name](prec) if name != "normal_small" else kabc.workloads.normal(prec, 100)
    s = kabc.SmcSession(ctx, prior, cost, kabc.smc_config(nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_retrys=retrys, mcmc_tol=0.3,
                                                        max_iterations=iters))
    s.init()
    stops = []
    for _ in range(iters):
        stops.append(s.iterate())
        if stops[-1]:
            break
    th, X, lpi, alive = s.state()
    sc = s.scalars()
    s.close()
    res = kabc.smc(prior, cost, nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_tol=0.3, max_iterations=iters, ctx=ctx)
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
        assert (z["th"].view(np.uint64) == th.view(np.uint64)).all(), f"theta differs on rank {rank}"
        assert (z["X"].view(np.uint64) == X.view(np.uint64)).all() and (z["lpi"].view(np.uint64) == lpi.view(np.uint64)).all()
        assert (z["alive"] == alive).all() and float(z["eps"]) == sc["eps"]
        assert int(z["evals"]) == sc["cost_evals"] and int(z["accepted"]) == sc["accepted"] and int(z["events"]) == sc["events"]
        assert list(z["stops"]) == stops
        r = np.load(os.path.join(str(tmp_path), f"run{rank}.npz"))
        assert (r["C"].view(np.uint64) == res.C.view(np.uint64)).all() and float(r["eps"]) == res.eps
        assert int(r["it"]) == res.iterations and int(r["evals"]) == res.cost_evals
        p0 = np.asarray(res.P[0].particles if len(prior) > 1 else res.P.particles)
        assert (r["P0"].view(np.uint64) == p0.view(np.uint64)).all()


# N is a multiple of 2, 3, 4 and 8.  retrys = 1: retry sweeps, launched kernel by kernel; retrys = 0: the iteration is replayed
# from a CUDA graph.  normal_small / lv_smc in F64 (bit-exact simulators), normal_smc with the F32 simulator.
SMC_CASES = [("normal_small", "f64", 6144, 8, 0), ("normal_small", "f64", 6144, 8, 1), ("normal_smc", "f32", 49152, 6, 0),
             ("lv_smc", "f64", 768, 4, 0), ("gk_ais", "f64", 192, 3, 0), ("ma2_smc", "f32", 24576, 5, 1)]


@pytest.mark.parametrize("name,prec,N,iters,retrys", SMC_CASES)
@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_multi_rank_smc_equals_single_rank(kabc, ctx, tmp_path, backend, name, prec, N, iters, retrys):
    worlds = _worlds(backend)
    if not worlds:
        pytest.skip("needs >= 2 GPUs")
    if backend == "gloo":
        worlds = worlds[:1] if (name, retrys) != ("normal_small", 0) else worlds  # shared-GPU ranks are slow: one world size
    for world in worlds:
        _launch(tmp_path, world, "smc", name, prec, N, iters, retrys, backend)
        _check_smc(kabc, ctx, tmp_path, world, name, prec, N, iters, retrys)


@pytest.mark.parametrize("name,prec,N,sweeps", [("normal_small", "f64", 1536, 4), ("gk_ais", "f32", 384, 2)])
@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_multi_rank_ais_equals_single_rank(kabc, ctx, tmp_path, backend, name, prec, N, sweeps):
    """Row e2: each rank moves N/2G red + N/2G black walkers; the moved colour reaches every replica through peer stores."""
    worlds = _worlds(backend)
    if not worlds:
        pytest.skip("needs >= 2 GPUs")
    if backend == "gloo":
        worlds = worlds[:1]
    prior, cost = kabc.workloads.WORKLOADS[name](prec) if name != "normal_small" else kabc.workloads.normal(prec, 100)
    a = kabc.AisSession(ctx, prior, cost, kabc.ais_config(N, 1, scale=0.5 if name == "gk_ais" else 0.05))
    a.init()
    a.sweep(sweeps)
    th, lp, ll = a.state()
    cn = a.counters()
    a.close()
    for world in worlds:
        _launch(tmp_path, world, "ais", name, prec, N, sweeps, 0, backend)
        for rank in range(world):
            z = np.load(os.path.join(str(tmp_path), f"rank{rank}.npz"))
            assert (z["th"].view(np.uint64) == th.view(np.uint64)).all(), f"theta differs on rank {rank} of {world}"
            assert (z["lp"].view(np.uint64) == lp.view(np.uint64)).all() and (z["ll"].view(np.uint64) == ll.view(np.uint64)).all()
            assert int(z["evals"]) == cn["cost_evals"] and int(z["accepted"]) == cn["accepted"]
