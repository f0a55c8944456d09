python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_eval.py normal_smc
for u in 1 4 8; do python scripts/bench_eval.py normal_smc build/variants/libkabc_unroll$u.so; done
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
prior, cost = k.workloads.normal("f32")
k.smc(prior, cost, nparticles=1<<12, epstol=0.0111, ctx=ctx)
for rep in range(3):
    t=time.perf_counter(); r=k.smc(prior, cost, nparticles=1<<20, epstol=0.0111, ctx=ctx); dt=time.perf_counter()-t
    print("smc() run:", dt, r.iterations, r.cost_evals/dt)
s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<20, epstol=0.0111))
t=time.perf_counter(); s.init(); n=0
while not s.iterate(): n+=1
dt=time.perf_counter()-t; print("stepwise:", dt, n+1)
PY
