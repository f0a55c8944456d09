# gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [pytest -k expr]'   : multi-GPU parity tests (log kept for profiles/), bench
# lines at 1 and N GPUs (the driver's own --steps 20 --warmup 5 window), warm per-kernel times on N GPUs, both sweep variants
N=${1:-2}
KEXPR=${2:-nccl}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${N}.txt
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -v --timeout 300 -x -k "$KEXPR" 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed|Error|assert" > gpurun_out/multi_parity_${N}gpu.txt
tail -3 gpurun_out/multi_parity_${N}gpu.txt
for V in queued split; do
if [ $V = split ]; then export KABC_SWEEP=split; else unset KABC_SWEEP; fi
for G in 1 $N; do
  if [ $G = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29618"; fi
  timeout 300 $L bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>gpurun_out/bench_${G}gpu_$V.err | grep '^{' > gpurun_out/bench_${G}gpu_$V.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${G}gpu_$V.json'));print('bench $V',d['n_gpus'],'%.4g'%d['value'],'%.4f'%d['ms_per_step'],'e2e %.4g'%d.get('e2e',{}).get('value',0),d.get('smc_time_to_eps_s'),d['kernel_times_us'],d['guard']['ok'])" || tail -5 gpurun_out/bench_${G}gpu_$V.err
done
for W in normal_smc ma2_smc; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py $W 2>/dev/null | grep world
done
done
