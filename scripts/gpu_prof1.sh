set -x
# launch list (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:k_smc_simulate -s 3 -c 2 -o gpurun_out/prof_sim_r1 \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_normal.json 2> gpurun_out/bench_normal.err; cat gpurun_out/bench_normal.json
