# memcheck + racecheck of the smc / AIS / ABCDE / pfilter paths at small sizes
# gpurun -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
cat > gpurun_out/san.py <<'PY'
import os, sys
sys.path.insert(0, '.')
import kissabc_jl_b200 as k
ctx = k.Context()
quick = os.environ.get("SAN_QUICK") == "1"
if not quick:
    for wl in ("normal_smc", "ma2_smc", "lv_smc", "gk_ais"):
        prior, cost = k.workloads.WORKLOADS[wl]("f32") if wl != "gk_ais" else k.workloads.gk("f32", 1000)
        if wl == "lv_smc": prior, cost = k.workloads.lv("f32", cap=2000)
        r = k.smc(prior, cost, nparticles=2000 if wl != "gk_ais" else 200, alpha=0.8, min_r_ess=0.5, max_iterations=4, ctx=ctx)
        print(wl, "smc", r.eps, r.iterations)
        post = k.ApproxKernelizedPosterior(prior, cost, 0.5)
        out = k.sample(post, k.AIS(64), 128, ntransitions=2, ctx=ctx)
        print(wl, "ais", out[0].mean())
# discrete priors + hard-threshold posterior + the two other samplers (small and > 4096 particles: both sort paths)
R = -30.0 ** 2 / (30.0 - 15.0 ** 2)
pri = k.Factored(k.NegativeBinomial(R, R / (30.0 + R)), k.Beta(15, 2))
r = k.smc(pri, k.Socks((0, 11), 11), nparticles=1500, alpha=0.9, max_iterations=5, ctx=ctx)
print("socks smc", r.eps, r.iterations)
out = k.sample(k.ApproxPosterior(pri, k.Socks((0, 11), 11), 0.1), k.AIS(64), 128, ntransitions=2, ctx=ctx)
print("socks ais", out[0].mean())
prior, cost = k.workloads.normal("f32", 100)
for n in (300, 5000):
    a = k.ABCDE(prior, cost, 0.05, nparticles=n, generations=5, alpha=0.3, ctx=ctx)
    f = k.pfilter(prior, cost, n, max_iters=3, ctx=ctx)
    print("abcde/pfilter", n, a.nsim, f.eps, f.nreps)
print("done")
PY
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python gpurun_out/san.py > gpurun_out/san_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/san_memcheck.log
tail -12 gpurun_out/san_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 3 python gpurun_out/san.py > gpurun_out/san_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/san_racecheck.log
tail -12 gpurun_out/san_racecheck.log
