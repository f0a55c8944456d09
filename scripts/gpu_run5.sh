set -x
for w in ma2_smc lv_smc gk_ais; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
timeout 600 python bench.py --precision f64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_normal_f64.json 2>&1; cat gpurun_out/bench_normal_f64.json
ncu --set full --clock-control none --import-source on -k regex:"k_smc_propose|k_sel_final|k_sel_hist|k_resample_gather" -s 12 -c 5 -o gpurun_out/prof_ctrl_r1 \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ctrl.log 2>&1
tail -2 gpurun_out/ncu_ctrl.log
