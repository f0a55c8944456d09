set -x
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for g in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_${g}gpu_p2p.json; python -c "import json;d=json.load(open('gpurun_out/bench_${g}gpu_p2p.json'));print('P2P',d['n_gpus'],d['value'],d['ms_per_step'])"
  KABC_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2962$g bench.py --gpus $g --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_${g}gpu_nccl.json; python -c "import json;d=json.load(open('gpurun_out/bench_${g}gpu_nccl.json'));print('NCCL',d['n_gpus'],d['value'],d['ms_per_step'])"
done
python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_1gpu_b.json; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_b.json'));print(d['n_gpus'],d['value'],d['ms_per_step'])"
