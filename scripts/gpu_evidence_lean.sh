# lean evidence refresh (about 70 s): full ncu capture of the hot kernel, launch list, headline bench line
ncu --set full --clock-control none --import-source on -k regex:"k_smc_simulate" -s 10 -c 1 -o gpurun_out/prof_final2 \
    python bench.py --steps 12 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_final2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/launches_final2.csv \
    python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches2.log 2>&1
python bench.py > gpurun_out/BENCH_normal_smc2.json 2>/dev/null; cat gpurun_out/BENCH_normal_smc2.json | cut -c1-300
