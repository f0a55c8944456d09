# 2 GPUs: packed inbox pushes vs direct row pushes
KABC_PACKED_PUSH=1 timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
for pk in 1 0; do
KABC_PACKED_PUSH=$pk timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$pk bench.py --gpus 2 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print('packed=$pk',d['value'],d['ms_per_step'])"
KABC_PACKED_PUSH=$pk timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2963$pk scripts/multi_profile.py normal_smc 2>/dev/null | grep world
done
