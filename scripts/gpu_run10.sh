set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python - <<'PY'
import sys
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
for wl, prec in (("normal_smc","f32"),("ma2_smc","f32")):
    prior, cost = k.workloads.WORKLOADS[wl](prec)
    s = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=1<<20))
    s.init(); s.iterate_n(30, ignore_stop=True)
    acc = {}
    for _ in range(10):
        for kk,v in s.profile_iteration().items(): acc[kk] = acc.get(kk,0)+v/10
    print(wl, {kk: round(v,1) for kk,v in acc.items()}, "sum", round(sum(acc.values()),1))
PY
python bench.py --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_1gpu_d.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_d.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
for g in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_${g}gpu_p2p3.json; python -c "import json;d=json.load(open('gpurun_out/bench_${g}gpu_p2p3.json'));print('P2P',d['n_gpus'],d['value'],d['ms_per_step'])"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2963$g scripts/multi_profile.py normal_smc 2>/dev/null | grep world
done
