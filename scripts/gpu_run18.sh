python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
for g in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_${g}gpu_shard.json; python -c "import json;d=json.load(open('gpurun_out/bench_${g}gpu_shard.json'));print('SHARD',d['n_gpus'],d['value'],d['ms_per_step'])"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2963$g scripts/multi_profile.py normal_smc 2>/dev/null | grep world
done
KABC_SHARD_ROWS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29654 bench.py --gpus 4 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_4gpu_repl.json; python -c "import json;d=json.load(open('gpurun_out/bench_4gpu_repl.json'));print('REPL',d['n_gpus'],d['value'],d['ms_per_step'])"
python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_1gpu_g.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_g.json'));print(d['n_gpus'],d['value'],d['ms_per_step'])"
