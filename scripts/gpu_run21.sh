python - <<'PY'
import sys
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
prior, cost = k.workloads.WORKLOADS["normal_smc"]("f32")
s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<20))
s.init(); s.iterate_n(30, ignore_stop=True)
print("first profiled (prefetched variates available):", {kk: round(v,1) for kk,v in s.profile_iteration().items()})
print("second profiled (inline variates):", {kk: round(v,1) for kk,v in s.profile_iteration().items()})
import time
for n in (50,):
    d, ms = s.iterate_n(n, ignore_stop=True); print("iterate_n", n, ms/n, "ms/iter")
PY
