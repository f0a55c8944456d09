# 1 GPU: parity suite, then the queued sweep with different bounds on the proposal tiles in flight (KABC_PROP_CAP)
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -30 gpurun_out/smoke.txt; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 240 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
for CAP in 0 148 296 592; do
  export KABC_PROP_CAP=$CAP
  for W in normal_smc ma2_smc; do
  timeout 200 python bench.py --workload $W --no-cpu-baseline --steps 20 --no-extra --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cap $CAP', d['config']['workload'], '%.4g' % d['value'], '%.4f' % d['ms_per_step'], d['kernel_times_us'])"
  done
done
unset KABC_PROP_CAP
timeout 100 python scripts/kernel_times.py null_smc
