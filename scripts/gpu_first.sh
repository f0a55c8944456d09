set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python - <<'PY'
import sys, time
sys.path.insert(0,'.')
import kissabc_jl_b200 as k
ctx = k.Context()
names = ["FFMA","IMAD","IMAD.WIDE(+IADD)","LOP3","MUFU.LG2","MUFU.SIN","MUFU.SQRT","DFMA","I2F(+LOP)","philox words","f32 normals","f64 normals"]
import json
out = {}
for kind,n in enumerate(names):
    r, ms = ctx.microbench(kind)
    out[n] = r
    print(f"{n:20s} {r:.4e} ops/s  ({ms:.3f} ms)  per SM per clk @1.9GHz: {r/148/1.9e9:.1f}")
json.dump(out, open('gpurun_out/microbench.json','w'), indent=1)
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
