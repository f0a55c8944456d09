# gpurun -- 'bash scripts/gpu_tests.sh'   : GPU parity suite + one bench line + warm per-kernel times of the headline workload
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --no-cpu-baseline --steps 20 2>gpurun_out/bench.err | grep '^{' > gpurun_out/bench_1gpu.json
python -c "import json;d=json.load(open('gpurun_out/bench_1gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('smc_time_to_eps_s'),d['gpu_launches'])"
tail -3 gpurun_out/bench.err
timeout 300 python scripts/kernel_times.py normal_smc ma2_smc
