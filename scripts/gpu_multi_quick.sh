# gpurun --gpus N -- 'bash scripts/gpu_multi_quick.sh N' : smoke, nccl parity subset, bench at 1 and N GPUs, per-kernel times at N
N=${1:-2}
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
KABC_TEST_WORLDS=${KABC_TEST_WORLDS:-$N} timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -v --timeout 300 -x -k "nccl and (normal_small or lv_smc or ais or normal_smc)" 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed|Error|assert" > gpurun_out/multi_parity_${N}gpu.txt
tail -4 gpurun_out/multi_parity_${N}gpu.txt
for G in 1 $N; do
  if [ $G = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29618"; fi
  timeout 300 $L bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>gpurun_out/bench_${G}gpu.err | grep '^{' > gpurun_out/bench_${G}gpu.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${G}gpu.json'));print('bench',d['n_gpus'],'%.4g'%d['value'],'%.4f'%d['ms_per_step'],'e2e %.4g'%d.get('e2e',{}).get('value',0),d.get('smc_time_to_eps_s'),d['kernel_times_us'],d['guard']['ok'])" || tail -5 gpurun_out/bench_${G}gpu.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
