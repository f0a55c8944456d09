# gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'   : multi-GPU parity tests (log kept for profiles/), bench lines at 1 and N
# GPUs (the driver's own --steps 20 --warmup 5 window) and warm per-kernel times on N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${N}.txt
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v --timeout 300 -x 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed|Error|assert" > gpurun_out/multi_parity_${N}gpu.txt
tail -4 gpurun_out/multi_parity_${N}gpu.txt
for G in 1 $N; do
  if [ $G = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29618"; fi
  timeout 300 $L bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>gpurun_out/bench_${G}gpu.err | grep '^{' > gpurun_out/bench_${G}gpu.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${G}gpu.json'));print('bench',d['n_gpus'],d['value'],d['ms_per_step'],d.get('e2e',{}).get('value'),d.get('smc_time_to_eps_s'),d['kernel_times_us'],d['per_rank_ms_per_step'],d['guard'])" || tail -5 gpurun_out/bench_${G}gpu.err
done
for W in normal_smc ma2_smc; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py $W 2>/dev/null | grep world
done
