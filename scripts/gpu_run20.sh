python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline 2>gpurun_out/b1.err | grep '^{' > gpurun_out/bench_1gpu_i.json; tail -3 gpurun_out/b1.err; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_i.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['gpu_launches'])"
python bench.py --workload ma2_smc --no-cpu-baseline --steps 50 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print('ma2',d['value'],d['ms_per_step'],d['e2e']['value'])"
