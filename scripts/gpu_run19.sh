python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline 2>gpurun_out/b1.err | grep '^{' > gpurun_out/bench_1gpu_h.json; tail -3 gpurun_out/b1.err; python -c "import json;d=json.load(open('gpurun_out/bench_1gpu_h.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['clocks'])"
for g in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2961$g bench.py --gpus $g --no-cpu-baseline 2>gpurun_out/b$g.err | grep '^{' > gpurun_out/bench_${g}gpu_h.json; tail -3 gpurun_out/b$g.err; python -c "import json;d=json.load(open('gpurun_out/bench_${g}gpu_h.json'));print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['smc_time_to_eps_s'],d['clocks'])"
done
python -c "import __graft_entry__ as g; g.smoke()"
