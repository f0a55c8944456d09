python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/bench_eval.py normal_smc
for u in 1 4 8; do python scripts/bench_eval.py normal_smc build/variants/libkabc_unroll$u.so; done
python bench.py --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_1gpu_f.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_f.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
