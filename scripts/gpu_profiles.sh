# gpurun -- 'bash scripts/gpu_profiles.sh'   : round-end evidence on ONE GPU: ncu --set full captures of the dominant kernel of
# every workload (-> profiles/instr_table.json via scripts/ncu_instr_table.py, run on the CPU box), ncu launch list of the bench
# command, warm / cold per-kernel times, the full bench line (extras + CPU baseline) and the reference arm.
mkdir -p gpurun_out
if ! timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; then echo "SMOKE FAILED"; tail -20 gpurun_out/smoke.txt; exit 1; fi
cap() { # workload kernel-regex skipped-launches iterations
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 \
      python scripts/ncu_target.py $1 $4 > gpurun_out/ncu_$1.log 2>&1
  ls -la gpurun_out/prof_$1.ncu-rep
}
cap normal_smc k_smc_sweep_q 10 12
cap ma2_smc k_smc_sweep_q 10 12
cap lv_smc k_smc_simulate_lv 3 5
cap gk_ais k_ais_simulate_gk 2 3
KABC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra --no-guard > gpurun_out/ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_r2.csv | tee gpurun_out/launches_r2.txt | head -12
timeout 200 python scripts/kernel_times.py normal_smc ma2_smc lv_smc
timeout 200 python scripts/kernel_times.py normal_smc --cold
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/BENCH_normal_smc.json 2> gpurun_out/BENCH_normal_smc.err; head -c 1500 gpurun_out/BENCH_normal_smc.json; echo
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/BENCH_reference.json 2> gpurun_out/BENCH_reference.err; head -c 600 gpurun_out/BENCH_reference.json; echo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
