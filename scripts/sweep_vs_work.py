"""Sweep-kernel time against the number of cost evaluations of the sweep, over population sizes around 2^20 (one GPU):
is the persistent sweep's time a staircase in ceil(evaluations / resident threads)?  usage: python scripts/sweep_vs_work.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kissabc_jl_b200 as k  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "normal_smc"
ctx = k.Context()
slots = ctx.sm_count() * 6 * 256
prior, cost = k.workloads.WORKLOADS[wl]("f32")
rows = []
for f in [0.55, 0.62, 0.68, 0.74, 0.80, 0.86, 0.92, 0.97, 1.0, 1.03, 1.08, 1.14, 1.20, 1.26, 1.32, 1.40]:
    n = int((1 << 20) * f) & ~255
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=n))
    s.init()
    s.iterate_n(30, ignore_stop=True)
    t, ev = 0.0, 0
    for _ in range(6):
        e0 = s.scalars()["cost_evals"]
        t += s.profile_iteration()["sweep"] / 6
        ev += (s.scalars()["cost_evals"] - e0) / 6
    rows.append(dict(n=n, evals=ev, rounds=ev / slots, sweep_us=round(t, 1), ns_per_eval=round(1e3 * t / ev, 4)))
    print(rows[-1], flush=True)
    s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(dict(workload=wl, slots=slots, rows=rows), open(os.path.join(ROOT, "gpurun_out", "sweep_vs_work_%s.json" % wl), "w"), indent=1)
