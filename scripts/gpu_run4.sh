set -x
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --no-cpu-baseline > gpurun_out/bench_normal_v2.json 2> gpurun_out/bench_v2.err; tail -3 gpurun_out/bench_v2.err; cat gpurun_out/bench_normal_v2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json
