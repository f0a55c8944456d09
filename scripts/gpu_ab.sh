# A/B of the sweep variants (1 GPU): lanes per particle of the normal simulator, fused sweep
mkdir -p gpurun_out
for L in 4 1 2 8; do echo "== work list, $L lanes/particle"; KABC_SIM_LANES=$L timeout 120 python scripts/kernel_times.py normal_smc; done
echo "== ma2 work list"; timeout 120 python scripts/kernel_times.py ma2_smc
echo "== fused sweep"; KABC_FUSED=1 timeout 120 python scripts/kernel_times.py normal_smc ma2_smc
for L in 4 1; do echo "== bench, $L lanes"; KABC_SIM_LANES=$L timeout 200 python bench.py --no-cpu-baseline --steps 20 --no-extra --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_times_us'])"; done
