# 4 GPUs, packed inbox pushes
KABC_PACKED_PUSH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 100 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys;d=json.loads(sys.stdin.read());print('packed=1',d['value'],d['ms_per_step'])"
KABC_PACKED_PUSH=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 scripts/multi_profile.py normal_smc 2>/dev/null | grep world
