# gpurun -- 'bash scripts/gpu_tests.sh'   : GPU parity suite + one bench line + warm per-kernel times of the headline workload
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | grep '^{' > gpurun_out/bench_1gpu.json
python -c "import json;d=json.load(open('gpurun_out/bench_1gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['gpu_launches'])"
python scripts/kernel_times.py normal_smc ma2_smc
