python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --no-cpu-baseline 2>gpurun_out/b1.err | grep '^{' > gpurun_out/bench_1gpu_k.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_k.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['gpu_launches'])"
KABC_NO_GRAPH=1 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print('nograph',d['value'],d['ms_per_step'])"
for w in ma2_smc lv_smc; do python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['config']['workload'],d['value'],d['ms_per_step'],d['e2e']['value'])"; done
