"""Worker of tests/test_gpu_multi.py: one rank of a multi-rank smc / AIS run (launched by torch.distributed.run).

backend nccl: one GPU per rank, the context moves its arena handles through NCCL.
backend gloo: the arena handles travel through torch.distributed; ranks may SHARE a GPU (rank % device_count), which is
how the multi-rank path is exercised on a 1-GPU box (the driver time-slices the ranks' kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import kissabc_jl_b200 as k
    from common import SEED
    out_dir, mode, name, prec, N, iters, retrys, backend = (sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]),
                                                            int(sys.argv[6]), int(sys.argv[7]), sys.argv[8])
    rank, world, local = k.dist.env_rank_world()
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    prior, cost = k.workloads.WORKLOADS[name](prec) if name != "normal_small" else k.workloads.normal(prec, 100)
    d = len(prior)
    L = k.lib()
    arena = max(L.kabc_smc_arena_bytes(N, d, world), L.kabc_ais_arena_bytes(N, d, world))
    ctx = k.dist.make_context(SEED, device_index=dev, arena_bytes=arena)
    if mode == "smc":
        s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_retrys=retrys, mcmc_tol=0.3,
                                                        max_iterations=iters))
        s.init()
        stops = []
        for _ in range(iters):
            stops.append(s.iterate())
            if stops[-1]:
                break
        th, X, lpi, alive = s.state()
        sc = s.scalars()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), th=th, X=X, lpi=lpi, alive=alive, eps=sc["eps"], evals=sc["cost_evals"],
                 events=sc["events"], accepted=sc["accepted"], stops=np.array(stops))
        s.close()
        # the one-call entry point as well (look-ahead queue + gather of the result on every rank)
        res = k.smc(prior, cost, nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_tol=0.3, max_iterations=iters, ctx=ctx)
        np.savez(os.path.join(out_dir, f"run{rank}.npz"), C=res.C, eps=res.eps, it=res.iterations, evals=res.cost_evals,
                 P0=np.asarray(res.P[0].particles if d > 1 else res.P.particles))
    else:
        a = k.AisSession(ctx, prior, cost, k.ais_config(N, 1, scale=0.5 if name == "gk_ais" else 0.05))
        a.init()
        a.sweep(iters)
        th, lp, ll = a.state()
        cn = a.counters()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), th=th, lp=lp, ll=ll, evals=cn["cost_evals"], accepted=cn["accepted"])
        a.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
