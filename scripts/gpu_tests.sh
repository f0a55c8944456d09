# gpurun -- 'bash scripts/gpu_tests.sh [pytest args]' : smoke first (a hang costs 90 s, not the whole budget), then the GPU parity
# suite with a per-test timeout and a streamed log, one bench line and warm per-kernel times of the headline workload
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then
  echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.txt; exit 1
fi
tail -2 gpurun_out/smoke.txt
timeout ${PYTEST_BUDGET:-1200} python -m pytest tests -m gpu -x -v --timeout 240 --durations=15 "$@" > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu.txt | tail -5
grep -E "FAILED|ERROR|Timeout" gpurun_out/pytest_gpu.txt | head -10
timeout 240 python bench.py --no-cpu-baseline --steps 20 --no-extra 2>gpurun_out/bench.err | grep '^{' > gpurun_out/bench_1gpu.json
python -c "import json;d=json.load(open('gpurun_out/bench_1gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('smc_time_to_eps_s'),d['gpu_launches'],d['kernel_times_us'],d['roofline'].get('frac'))"
tail -3 gpurun_out/bench.err
timeout 200 python scripts/kernel_times.py normal_smc ma2_smc
