# gpurun -- 'bash scripts/gpu_profiles.sh'   : round-end evidence: bench lines for every workload, ncu launch list, one
# `--set full` capture of the two top kernels, warm and cold (L2 flushed) per-kernel times.  Copy what matters to profiles/.
set -x
python bench.py > gpurun_out/BENCH_normal_smc.json 2> gpurun_out/BENCH_normal_smc.err; cat gpurun_out/BENCH_normal_smc.json
for w in ma2_smc lv_smc gk_ais; do
  timeout 900 python bench.py --workload $w --steps 30 --warmup 3 $EXTRA_BENCH_FLAGS > gpurun_out/BENCH_$w.json 2>/dev/null
done
python bench.py --precision f64 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/BENCH_normal_smc_f64.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_smc_simulate|k_smc_propose" -s 20 -c 4 -o gpurun_out/prof_final \
    python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
python scripts/kernel_times.py normal_smc ma2_smc lv_smc
python scripts/kernel_times.py normal_smc --cold
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
