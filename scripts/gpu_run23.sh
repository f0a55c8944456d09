python -m pytest tests -m gpu -q 2>&1 | tail -40
python - <<'PY'
import sys
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
for wl in ("normal_smc","ma2_smc"):
    prior, cost = k.workloads.WORKLOADS[wl]("f32")
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<20))
    s.init(); s.iterate_n(30, ignore_stop=True)
    acc = {}
    for _ in range(10):
        for kk,v in s.profile_iteration().items(): acc[kk] = acc.get(kk,0)+v/10
    print(wl, {kk: round(v,1) for kk,v in acc.items()}, "sum", round(sum(acc.values()),1))
PY
python bench.py --no-cpu-baseline 2>gpurun_out/b1.err | grep '^{' > gpurun_out/bench_1gpu_j.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_j.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
