# gpurun --gpus N -- 'bash scripts/gpu_ab2.sh N' : queued-sweep variants on N GPUs (tile size x prefetch)
N=${1:-2}
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618"
for V in "256 0" "512 1" "256 1" "512 0"; do
  set -- $V
  export KABC_TILE=$1 KABC_PREFETCH=$2
  timeout 300 $L bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-e2e 2>gpurun_out/bench_ab.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tile $1 prefetch $2 :', d['n_gpus'], '%.4g' % d['value'], '%.4f' % d['ms_per_step'], d['kernel_times_us'])" || tail -3 gpurun_out/bench_ab.err
done
