"""Target of the ncu captures: init + K iterations of one workload at full size, launched kernel by kernel (KABC_NO_GRAPH=1 is
set here so that every kernel is its own launch); writes the units every iteration processed to gpurun_out/ncu_units_<wl>.json
so that scripts/ncu_instr_table.py can divide the executed instructions of the captured launch by them.
usage: python scripts/ncu_target.py <workload> [iterations] [precision]"""
import json
import os
import sys

os.environ["KABC_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kissabc_jl_b200 as k  # noqa: E402

wl = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 14
prec = sys.argv[3] if len(sys.argv) > 3 else "f32"
ctx = k.Context(seed=0x4B49535341424300)
prior, cost = k.workloads.WORKLOADS[wl](prec)
units = []
if wl == "gk_ais":
    a = k.AisSession(ctx, prior, cost, k.ais_config(1 << 18, 1, scale=0.5))
    a.init()
    e0 = a.counters()["cost_evals"]
    for _ in range(iters):
        a.sweep(1)
        e1 = a.counters()["cost_evals"]
        units.append({"evals": e1 - e0, "launches_per_step": 2})  # two half-steps per sweep
        e0 = e1
else:
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1 << 20, epstol=0.0))
    s.init()
    sc0 = s.scalars()
    for _ in range(iters):
        s.iterate_n(1, ignore_stop=True)
        sc = s.scalars()
        units.append({"evals": sc["cost_evals"] - sc0["cost_evals"], "events": sc["events"] - sc0["events"], "launches_per_step": 1})
        sc0 = sc
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"workload": wl, "precision": prec, "units": units}, open(os.path.join(ROOT, "gpurun_out", f"ncu_units_{wl}.json"), "w"))
print(wl, units[-1])
