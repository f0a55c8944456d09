"""Kernel-only throughput of kabc_eval_cost_device for each registered simulator (theta resident in HBM)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import kissabc_jl_b200 as k
if len(sys.argv) > 2:
    k._capi.LIB_PATH = sys.argv[2]; k._capi._lib = None
which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = k.Context()
L = ctx.L
def run(name, prec, n, reps=3):
    prior, cost = k.workloads.WORKLOADS[name](prec)
    d = len(prior)
    th = ctx.prior_sample(prior, n)
    dth = torch.from_numpy(th).cuda().contiguous(); out = torch.empty(n, dtype=torch.float64, device="cuda")
    m = cost._pod(); ms = C.c_float(); best = 1e9
    for r in range(reps):
        k._capi.check(L.kabc_eval_cost_device(ctx.h, C.byref(m), d, C.c_void_p(dth.data_ptr()), n, 0, r, C.c_void_p(out.data_ptr()), C.byref(ms)))
        best = min(best, ms.value)
    print(f"{name:12s} {prec} n={n:8d} {best:9.3f} ms  {n/best*1e3:.4e} evals/s  finite={torch.isfinite(out).float().mean().item():.3f}", flush=True)
for name, n in (("normal_smc", 1 << 20), ("ma2_smc", 1 << 20), ("gk_ais", 1 << 16), ("lv_smc", 1 << 18)):
    if which in ("all", name):
        run(name, "f32", n)
