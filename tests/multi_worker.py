"""Worker of tests/test_gpu_multi.py: one rank of a multi-GPU smc run (launched by torch.distributed.run)."""
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
    out_dir, name, prec, N, iters = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    retrys = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    rank, world, local = k.dist.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = k.dist.make_context(SEED)
    prior, cost = k.workloads.WORKLOADS[name](prec) if name != "normal_small" else k.workloads.normal(prec, 100)
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=N, alpha=0.9, min_r_ess=0.7, mcmc_retrys=retrys, mcmc_tol=0.3, max_iterations=iters))
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
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
