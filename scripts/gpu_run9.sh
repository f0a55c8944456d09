for g in 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2963$g scripts/multi_profile.py normal_smc 2>/dev/null | grep world
KABC_NO_P2P=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2964$g scripts/multi_profile.py normal_smc 2>/dev/null | grep world
done
