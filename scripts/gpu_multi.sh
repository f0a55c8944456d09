# gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'   : multi-GPU parity tests, bench line and warm per-kernel times on N GPUs
N=${1:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --no-cpu-baseline 2>gpurun_out/bench_${N}gpu.err | grep '^{' > gpurun_out/bench_${N}gpu.json
python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
