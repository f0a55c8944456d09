ncu --set full --clock-control none --import-source on -k regex:k_eval_cost_gk -c 1 -o gpurun_out/prof_gk_r1 python scripts/bench_eval.py gk_ais > gpurun_out/ncu_gk.log 2>&1
tail -2 gpurun_out/ncu_gk.log
