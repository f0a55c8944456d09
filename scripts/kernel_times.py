"""Warm per-kernel times of one smc iteration (kabc_smc_profile_iteration), 2^20 particles, one GPU.
usage: python scripts/kernel_times.py [workload ...] [--lg 20] [--cold]   (--cold: L2 flushed before every profiled iteration)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import kissabc_jl_b200 as k  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
lg = int(sys.argv[sys.argv.index("--lg") + 1]) if "--lg" in sys.argv else 20
cold = "--cold" in sys.argv
ctx = k.Context()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if cold else None
out = {}
for wl in args or ["normal_smc"]:
    prior, cost = k.workloads.WORKLOADS[wl]("f32")
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1 << lg))
    s.init()
    s.iterate_n(30, ignore_stop=True)
    acc = {}
    for _ in range(10):
        if cold:
            flush.zero_()
            torch.cuda.synchronize()
        for kk, v in s.profile_iteration().items():
            acc[kk] = acc.get(kk, 0) + v / 10
    out[wl] = {kk: round(v, 1) for kk, v in acc.items()}
    print(wl, "cold" if cold else "warm", out[wl], "sum", round(sum(acc.values()), 1), flush=True)
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_times_%s.json" % ("cold" if cold else "warm")), "w"), indent=1)
