# how the per-iteration kernels grow with the population on ONE GPU (stand-in for the replicated state of a multi-GPU run)
cat > gpurun_out/prof_n.py <<'PY'
import sys, json
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
import os
for lg in [int(x) for x in os.environ.get("LGS","20,21,22,23").split(",")]:
    prior, cost = k.workloads.WORKLOADS["normal_smc"]("f32")
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<lg))
    s.init(); s.iterate_n(30, ignore_stop=True)
    acc = {}
    for _ in range(5):
        for kk,v in s.profile_iteration().items(): acc[kk] = acc.get(kk,0)+v/5
    print(lg, {kk: round(v,1) for kk,v in acc.items()}, "sum", round(sum(acc.values()),1), flush=True)
    s.close()
PY
python gpurun_out/prof_n.py
LGS=23 ncu --set full --clock-control none --import-source on -k regex:"k_smc_propose" -s 20 -c 2 -o gpurun_out/prof_propose_8m python gpurun_out/prof_n.py > gpurun_out/ncu_propose_8m.log 2>&1
tail -3 gpurun_out/ncu_propose_8m.log
