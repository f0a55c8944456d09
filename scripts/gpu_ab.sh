# A/B of the queued sweep with / without the cp.async prefetch stage (1 GPU)
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
for V in 1 0; do
  export KABC_PREFETCH=$V
  echo "== prefetch=$V"; timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
  timeout 200 python bench.py --no-cpu-baseline --steps 20 --no-extra --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['kernel_times_us'])"
done
