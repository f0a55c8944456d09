# gpurun --gpus N -- 'bash scripts/gpu_final_multi.sh N' : A/B of the proposal-tile bound at N GPUs, then parity + bench + per-kernel
# times with the better setting (the code default is then set to match; KABC_PROP_CAP and the default take the same path)
N=${1:-8}
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
: > gpurun_out/cap_ab_${N}gpu.txt
I=0
for CAP in 0 296 592 0; do
  I=$((I+1))
  KABC_PROP_CAP=$CAP timeout 75 $TR --master-port $((29600+I)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e --no-guard 2>>gpurun_out/bench_${N}gpu.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($CAP, '%.5f' % d['ms_per_step'], '%.4g' % d['value'], d['kernel_times_us'])" | tee -a gpurun_out/cap_ab_${N}gpu.txt
done
BEST=$(python -c "
import collections
t=collections.defaultdict(list)
for l in open('gpurun_out/cap_ab_${N}gpu.txt'):
    p=l.split(); t[p[0]].append(float(p[1]))
m={k:min(v) for k,v in t.items()}
b=min(m,key=m.get)
# keep the unbounded sweep unless a bound wins by more than 1 %
print(b if m[b] < 0.99*m.get('0',1e9) else 0)")
echo "best cap: $BEST" | tee -a gpurun_out/cap_ab_${N}gpu.txt
export KABC_PROP_CAP=$BEST
KABC_TEST_WORLDS=$N timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -v --timeout 120 -x -k "nccl and (normal_small or lv_smc or ais or normal_smc)" 2>&1 | grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed|Error|assert" > gpurun_out/multi_parity_${N}gpu.txt
tail -4 gpurun_out/multi_parity_${N}gpu.txt
for G in 1 $N; do
  if [ $G = 1 ]; then L="python"; else L="$TR --master-port 29618"; fi
  timeout 100 $L bench.py --gpus $G --steps 20 --warmup 5 --no-cpu-baseline --no-extra 2>gpurun_out/bench_${G}gpu.err | grep '^{' > gpurun_out/bench_${G}gpu.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${G}gpu.json'));print('bench',d['n_gpus'],'%.4g'%d['value'],'%.4f'%d['ms_per_step'],'e2e %.4g'%d.get('e2e',{}).get('value',0),d.get('smc_time_to_eps_s'),d['kernel_times_us'],d['guard']['ok'])" || tail -5 gpurun_out/bench_${G}gpu.err
done
timeout 75 $TR --master-port 29638 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
# the two other multi-GPU configurations of BASELINE.json: Lotka-Volterra smc (config 5) and g-and-k AIS (config 4)
for W in lv_smc gk_ais; do
  I=$((I+1))
  timeout 90 $TR --master-port $((29640+I)) bench.py --workload $W --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-extra --no-e2e --no-guard 2>gpurun_out/bench_${W}_${N}gpu.err | grep '^{' > gpurun_out/bench_${W}_${N}gpu.json
  python -c "import json;d=json.load(open('gpurun_out/bench_${W}_${N}gpu.json'));print('bench','$W',d['n_gpus'],'%.4g'%d['value'],'%.4f'%d['ms_per_step'],d.get('kernel_times_us'))" || tail -5 gpurun_out/bench_${W}_${N}gpu.err
done
