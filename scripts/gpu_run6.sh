set -x
python - <<'PY'
import sys, json
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
names = {9:"philox words",10:"f32 normals (I2FP)",12:"philox + 4 I2FP",13:"f32 normals (mantissa trick)",14:"philox + lg2,sqrt only"}
for kind,n in names.items():
    r, ms = ctx.microbench(kind)
    print(f"{n:32s} {r:.4e} /s  cycles per warp-block per SMSP @1.9GHz: {148*4*1.9e9*128/r:.1f}")
for wl, prec in (("normal_smc","f32"),("ma2_smc","f32")):
    prior, cost = k.workloads.WORKLOADS[wl](prec)
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<20))
    s.init()
    s.iterate_n(30, ignore_stop=True)
    acc = {}
    for _ in range(10):
        for kk,v in s.profile_iteration().items(): acc[kk] = acc.get(kk,0)+v/10
    print(wl, {kk: round(v,1) for kk,v in acc.items()}, "sum", round(sum(acc.values()),1))
PY
