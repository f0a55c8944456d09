# A/B of the two sweep forms on 1 GPU: queued (one persistent kernel) against split (propose kernel + simulate kernel)
mkdir -p gpurun_out
for V in queued split; do
  export KABC_SWEEP=$V
  echo "== sweep=$V"; timeout 120 python scripts/kernel_times.py normal_smc ma2_smc null_smc
  for W in normal_smc ma2_smc; do
  timeout 200 python bench.py --workload $W --no-cpu-baseline --steps 20 --no-extra --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['config']['workload'], d['value'], d['ms_per_step'], d['kernel_times_us'])"
  done
done
