set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_normal.json 2> gpurun_out/bench_normal.err; tail -5 gpurun_out/bench_normal.err; cat gpurun_out/bench_normal.json
