set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in gk_ais lv_smc ma2_smc; do
timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/bench_${w}_v3.json; python -c "import json;d=json.load(open('gpurun_out/bench_${w}_v3.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d.get('ssa_events_per_s'))"
done
