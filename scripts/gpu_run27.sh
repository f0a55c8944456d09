# 4 GPUs: multi-GPU parity (graph + NCCL capture + deferred pushes), bench line, warm per-kernel times
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 4 --no-cpu-baseline 2>gpurun_out/b4.err | grep '^{' > gpurun_out/bench_4gpu_k.json; tail -3 gpurun_out/b4.err; python -c "import json;d=json.load(open('gpurun_out/bench_4gpu_k.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['gpu_launches'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29638 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
KABC_NO_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 4 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print('nograph',d['value'],d['ms_per_step'])"
