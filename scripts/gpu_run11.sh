set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --workload gk_ais --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_gk_v2.json; python -c "import json;d=json.load(open('gpurun_out/bench_gk_v2.json'));print('gk',d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'])"
python bench.py --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_1gpu_e.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_e.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
