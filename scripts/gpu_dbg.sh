python -m pytest tests/test_gpu_golden.py -x -q 2>&1 | grep -E "Error|error|assert|kabc" | head -20
python - <<'PY'
import sys
sys.path.insert(0,'.')
import kissabc_jl_b200 as k, numpy as np
ctx = k.Context()
s = k.SmcSession(ctx, k.Uniform(1.4999999, 1.5000001), k.Deterministic(1, 1.5), k.smc_config(nparticles=5000, alpha=0.5, max_iterations=30))
s.init()
for it in range(30):
    try:
        st = s.iterate()
    except Exception as e:
        print("iteration", it+1, "failed:", e); break
    sc = s.scalars(); print(it+1, sc['eps'], sc['n_alive'], sc['accepted'], st)
    if st: break
PY
