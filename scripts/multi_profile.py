"""torchrun worker: warm per-kernel timings of the multi-rank smc iteration (rank 0 prints)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import kissabc_jl_b200 as k
rank, world, local = k.dist.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = k.dist.make_context(0x4B49535341424300)
wl = sys.argv[1] if len(sys.argv) > 1 else "normal_smc"
prior, cost = k.workloads.WORKLOADS[wl]("f32")
s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=(1 << 20) * world))
s.init(); s.iterate_n(30, ignore_stop=True)
acc = {}
for _ in range(10):
    for kk, v in s.profile_iteration().items():
        acc[kk] = acc.get(kk, 0) + v / 10
if rank == 0:
    print(wl, "world", world, {kk: round(v, 1) for kk, v in acc.items()}, "sum", round(sum(acc.values()), 1), flush=True)
dist.barrier(); dist.destroy_process_group()
