KABC_DEBUG_TIMING=1 python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
prior, cost = k.workloads.normal("f32")
k.smc(prior, cost, nparticles=1<<12, epstol=0.0111, ctx=ctx)
for rep in range(4):
    t=time.perf_counter(); r=k.smc(prior, cost, nparticles=1<<20, epstol=0.0111, ctx=ctx); dt=time.perf_counter()-t
    print("smc() run:", dt, r.iterations, r.cost_evals/dt, flush=True)
PY
