# 2 GPUs: multi-GPU parity tests, bench line, per-kernel warm times
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 2 --no-cpu-baseline 2>gpurun_out/b2.err | grep '^{' > gpurun_out/bench_2gpu_k.json; tail -3 gpurun_out/b2.err; python -c "import json;d=json.load(open('gpurun_out/bench_2gpu_k.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
