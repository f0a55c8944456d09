# memcheck + racecheck of the smc / AIS paths at small sizes
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
for wl in ("normal_smc","ma2_smc","lv_smc","gk_ais"):
    prior, cost = k.workloads.WORKLOADS[wl]("f32") if wl != "gk_ais" else k.workloads.gk("f32", 1000)
    if wl == "lv_smc": prior, cost = k.workloads.lv("f32", cap=2000)
    r = k.smc(prior, cost, nparticles=2000 if wl != "gk_ais" else 200, alpha=0.8, min_r_ess=0.5, max_iterations=4, ctx=ctx)
    print(wl, "smc", r.eps, r.iterations)
    post = k.ApproxKernelizedPosterior(prior, cost, 0.5)
    out = k.sample(post, k.AIS(64), 128, ntransitions=2, ctx=ctx)
    print(wl, "ais", out[0].mean())
print("done")
PY
compute-sanitizer --tool memcheck --error-exitcode 3 python /tmp/san.py 2>&1 | tail -8
echo "memcheck rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 3 python /tmp/san.py 2>&1 | tail -8
echo "racecheck rc=$?"
