#!/usr/bin/env python
"""bench.py -- headline benchmark of the KissABC hot path on B200.

Metric (BASELINE.json): cost evaluations / second of `smc(prior, cost)` on the README normal model at 2^20
particles per GPU, plus smc time-to-epsilon.  A "step" is one body of smc's `while true` loop (src/smc.jl:131-198):
epsilon quantile -> alive cut -> resample -> propose -> simulate+distance -> accept, over the whole population.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload normal_smc|ma2_smc|gk_ais|lv_smc]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU implementation (the C
oracle restatement, all host threads) on a bounded sample of the same workload: Julia is not installed here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SEED = 0x4B49535341424300
# algorithmic thread-instructions per unit of work: SURVEY.md section 8(d), fixed table (do not re-tune per run)
W_INSTR = {"normal_smc": 2.7e4, "ma2_smc": 3.0e3, "gk_ais": 6.7e5, "lv_smc": 52.0}
# algorithmic state bytes per cost evaluation of an smc sweep at d parameters: 8(3d+1)+1 read, 8(d+2) written
STATE_BYTES = lambda d: 8 * (3 * d + 1) + 1 + 8 * (d + 2)  # noqa: E731
EPS_TARGET = {"normal_smc": 0.0111, "ma2_smc": 0.1, "lv_smc": None, "gk_ais": None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def traffic_bytes(name, prec):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_top_kernels_r1.txt); None for workloads without a capture."""
    if name == "normal_smc" and prec == "f32":
        return 41.03e6 + 13.58e6
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_objects(k, name, prec):
    prior, cost = k.workloads.WORKLOADS[name](prec)
    return prior, cost, len(prior)


def oracle_objects(O, name):
    from common import GK_TARGET, LV_TARGET_X, LV_TARGET_Y, MA2_TARGET
    if name == "normal_smc":
        return O.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)]), O.make_model(O.NORMAL_MEANSTD, 1000, (2.0, 0.04), (50.0,)), 2
    if name == "ma2_smc":
        return O.make_priors([("uniform", -2, 2), ("uniform", -1, 1)]), O.make_model(O.MA2_AUTOCOV, 100, MA2_TARGET), 2
    if name == "gk_ais":
        return O.make_priors([("uniform", 0, 10)] * 4), O.make_model(O.GK_OCTILE, 10000, GK_TARGET, (0.8,)), 4
    return (O.make_priors([("uniform", -2, 1), ("uniform", -7, -4), ("uniform", -2, 1)]),
            O.make_model(O.LV_SSA, 0, LV_TARGET_X + LV_TARGET_Y, (50, 100, 30, 16, 20000)), 3)


def cpu_baseline(name, budget_s=12.0, threads=None):
    """The oracle (CPU restatement of the reference algorithm, FP64, OpenMP over particles exactly where the
    reference threads: src/smc.jl:122,168) timed on a bounded sample of the workload."""
    from oracle import oracle as O
    O.build()
    threads = threads or os.cpu_count() or 1
    pri, mod, d = oracle_objects(O, name)
    if name == "gk_ais":
        nw, sweeps = 256, 0
        a = O.Ais(SEED, pri, mod, O.ais_config(nw, 1, scale=0.5), nthreads=threads)
        a.init()
        e0 = a.counters()["cost_evals"]
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget_s:
            a.sweep(); sweeps += 1
        dt = time.perf_counter() - t0
        evals = a.counters()["cost_evals"] - e0
        return dict(value=evals / dt, unit="cost evals/s", cores=threads, kind="port",
                    sample=f"oracle AIS({nw}) red/black sweeps of the g-and-k model for {dt:.1f}s ({sweeps} sweeps, {evals} evals), FP64, OpenMP")
    n = {"normal_smc": 1 << 14, "ma2_smc": 1 << 16, "lv_smc": 1 << 11}[name]
    s = O.Smc(SEED, pri, mod, O.smc_config(nparticles=n), nthreads=threads)
    t0 = time.perf_counter()
    s.init()
    its = 0
    while time.perf_counter() - t0 < budget_s:
        s.iterate(); its += 1
    dt = time.perf_counter() - t0
    evals = s.scalars()["cost_evals"]
    return dict(value=evals / dt, unit="cost evals/s", cores=threads, kind="port",
                sample=f"oracle smc, {n} particles, init + {its} iterations in {dt:.1f}s ({evals} evals), FP64, OpenMP x{threads}")


REF_PARTICLES = {"normal_smc": 1 << 14, "ma2_smc": 1 << 16, "lv_smc": 1 << 11, "gk_ais": 256}


def run_reference(args):
    """Reference arm: the reference algorithm's CPU implementation (C oracle, all host threads) on the SAME step
    definition as the device arm -- one smc iteration (or one AIS sweep) -- over a bounded sample of the workload
    (REF_PARTICLES particles instead of 2^20).  W warm-up steps, then exactly K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    name = args.workload
    pri, mod, d = oracle_objects(O, name)
    n = REF_PARTICLES[name]
    t_all = time.perf_counter()
    if name == "gk_ais":
        obj = O.Ais(SEED, pri, mod, O.ais_config(n, 1, scale=0.5), nthreads=threads)
        obj.init()
        step, evals_now = obj.sweep, lambda: obj.counters()["cost_evals"]
    else:
        obj = O.Smc(SEED, pri, mod, O.smc_config(nparticles=n), nthreads=threads)
        obj.init()
        step, evals_now = obj.iterate, lambda: obj.scalars()["cost_evals"]
    max_s = args.ref_budget * (args.steps + args.warmup) if args.ref_budget else 150.0
    for _ in range(args.warmup):
        step()
    e0, t0, done = evals_now(), time.perf_counter(), 0
    for _ in range(args.steps):
        step(); done += 1
        if time.perf_counter() - t0 > max_s:  # keep the whole run within a few minutes whatever K is
            break
    dt = time.perf_counter() - t0
    v = (evals_now() - e0) / dt
    base = dict(value=v, unit="cost evals/s", cores=threads, kind="port",
                sample=f"oracle, {n} particles/walkers, {done} timed steps in {dt:.1f}s, FP64, OpenMP x{threads}")
    out = {"impl": "reference", "metric": "cost evals/sec", "value": v, "unit": "cost evals/s", "n_gpus": args.gpus,
           "steps": done, "warmup": args.warmup, "ms_per_step": dt / max(done, 1) * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": name, "particles": n,
                      "note": "CPU restatement of the reference (C oracle), not Julia: julia is not installed"},
           "cpu_baseline": base, "e2e": {"value": v, "unit": "cost evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": time.perf_counter() - t_all}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="normal_smc", choices=list(W_INSTR))
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--log2-particles", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=0.0, help="seconds of CPU work per reference step (0 = auto)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import kissabc_jl_b200 as k

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    name = args.workload
    prior, cost, d = workload_objects(k, name, args.precision)
    log2n = args.log2_particles or (18 if name == "gk_ais" else 20)
    n_per_gpu = 1 << log2n
    N = n_per_gpu * world  # weak scaling: per-GPU work fixed

    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(k.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx = k.Context(device=local_rank, seed=SEED, rank=rank, world=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    else:
        ctx = k.Context(device=local_rank, seed=SEED)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    is_ais = name == "gk_ais"
    if is_ais:
        sess = k.AisSession(ctx, prior, cost, k.ais_config(N, 1, scale=0.5))
    else:
        sess = k.SmcSession(ctx, prior, cost, k.smc_config(nparticles=N, epstol=0.0))
    sess.init()

    def evals_now():
        return sess.counters()["cost_evals"] if is_ais else sess.scalars()["cost_evals"]

    def step():
        flush.zero_()  # L2 flush between timed iterations (outside the event-timed region)
        torch.cuda.synchronize()
        if is_ais:
            return sess.sweep(1)
        return sess.iterate_n(1, ignore_stop=True)[1]

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0, e0 = sess.kernel_launches(), evals_now()
    ev0 = 0 if is_ais else sess.scalars()["events"]
    t_wall0 = time.perf_counter()
    ms_total = 0.0
    for _ in range(args.steps):
        ms_total += step()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    evals = evals_now() - e0
    launches = sess.kernel_launches() - l0
    events = 0 if is_ais else sess.scalars()["events"] - ev0

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    # with the replicated-state design every rank reports the same global counter
    value = evals / (ms_total * 1e-3)

    # ---- end to end through the public API (host buffers in, host buffers out).  smc: a whole run to the target
    # epsilon, every rank calls it (the call is collective), timed on the host as the max over ranks.
    e2e = None
    if not args.no_e2e and not is_ais:
        import ctypes as C
        eps_t = EPS_TARGET[name]
        kw = dict(nparticles=N, ctx=ctx)
        if eps_t is not None:
            kw["epstol"] = eps_t
        else:
            kw["max_iterations"] = 30
        sess.close()  # give the buffers back to the context cache
        k.smc(prior, cost, **kw)  # warm the call path (allocations, lazy module load)
        barrier()
        t0 = time.perf_counter()
        res = k.smc(prior, cost, **kw)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = d * C.sizeof(k._capi.PriorT) + C.sizeof(k._capi.ModelT) + C.sizeof(k._capi.SmcConfigT)
        d2h = N * (8 * d + 1 + 8) + 56 * res.iterations
        e2e = {"value": res.cost_evals / dt, "unit": "cost evals/s", "h2d_bytes_per_step": h2d / max(res.iterations, 1),
               "d2h_bytes_per_step": d2h / max(res.iterations, 1),
               "call": "kissabc_jl_b200.smc(prior, cost, nparticles=N, epstol=target) on every rank",
               "iterations": res.iterations, "cost_evals": res.cost_evals, "eps": res.eps, "time_s": dt, "eps_target": eps_t}
    elif not args.no_e2e and world == 1 and is_ais:
        sess.close()
        post = k.ApproxKernelizedPosterior(prior, cost, 0.5)
        t0 = time.perf_counter()
        _, cnt = k.sample(post, k.AIS(N), N, ntransitions=2, ctx=ctx, return_counters=True)
        dt = time.perf_counter() - t0
        e2e = {"value": cnt["cost_evals"] / dt, "unit": "cost evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": N * 8 * d,
               "call": "kissabc_jl_b200.sample(ApproxKernelizedPosterior(...), AIS(N), N, ntransitions=2)", "time_s": dt}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks, peak_src = measured_peaks()
    sm_count = ctx.sm_count()
    f_clk = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    r_issue = sm_count * 4 * 32 * f_clk * world
    units = events if name == "lv_smc" else evals
    achieved_instr = units * W_INSTR[name] / (ms_total * 1e-3)
    hbm_bytes = evals * STATE_BYTES(d)
    roofline = {
        "bound": "issue", "kernel": "k_ais_simulate*" if is_ais else "k_smc_simulate<model,precision>",
        "achieved": achieved_instr / 1e12, "peak": r_issue / 1e12, "unit": "T thread-instr/s (SM issue slots: SMs x 4 x 32 x f_clk at the sampled clock)",
        "frac": achieved_instr / r_issue, "traffic": traffic_bytes(name, args.precision),
        "w_instr_per_unit": W_INSTR[name], "unit_of_work": "SSA event" if name == "lv_smc" else "cost eval",
        "hbm": {"achieved": hbm_bytes / (ms_total * 1e-3) / 1e9, "peak": peaks["hbm_gbs"] * world, "unit": "GB/s",
                "frac": hbm_bytes / (ms_total * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world), "peak_source": peak_src,
                "note": "state sweep bytes only: the path is instruction-bound, not HBM-bound (SURVEY.md 8d)"},
    }
    out = {
        "metric": "cost evals/sec", "value": value, "unit": "cost evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 simulator draws, f64 state/distance/accept" if args.precision == "f32" else "f64",
        "data": "synthetic",
        "config": {"workload": name, "particles_per_gpu": n_per_gpu, "particles_total": N, "precision": args.precision,
                   "l2": "flushed between timed steps (256 MiB write, outside the event-timed region)",
                   "step": "one AIS red/black sweep" if is_ais else "one smc iteration (quantile, cut, resample, propose, simulate, accept)"},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_s_timed_region": wall,
        "cost_evals_timed": int(evals),
    }
    if name == "lv_smc":
        out["ssa_events_per_s"] = events / (ms_total * 1e-3)

    if e2e is not None:
        out["e2e"] = e2e
        if "eps_target" in e2e:
            out["smc_time_to_eps_s"] = e2e["time_s"]
            out["eps_target"] = e2e["eps_target"]

    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(name)
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
