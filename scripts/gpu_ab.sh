# A/B of the sweep variants (1 GPU) + true kernel durations from an ncu launch list
mkdir -p gpurun_out
echo "== queued sweep (default)"; timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
echo "== split: propose + simulate"; KABC_SWEEP=split timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
for V in queued split; do
  if [ $V = split ]; then export KABC_SWEEP=split; else unset KABC_SWEEP; fi
  echo "== bench, $V"; timeout 200 python bench.py --no-cpu-baseline --steps 20 --no-extra --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_times_us'])"
done
unset KABC_SWEEP
KABC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 70 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline --no-extra --no-guard > gpurun_out/ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_r2.csv 2>/dev/null | head -30
