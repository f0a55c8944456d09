"""CPU tests (no GPU): the C-ABI library builds, loads and exports every symbol include/kissabc_cuda.h declares;
the host mirror validates arguments like the reference; the product never imports the oracle; no compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(kabc):
    hdr = open(os.path.join(ROOT, "include", "kissabc_cuda.h")).read()
    declared = set(re.findall(r"\b(kabc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"kabc_status"}
    assert declared == set(kabc.SYMBOLS), declared ^ set(kabc.SYMBOLS)
    L = C.CDLL(kabc.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), sym
    out = subprocess.run(["nm", "-D", "--defined-only", kabc.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (kabc_[a-z0-9_]+)", out))
    assert declared <= exported


def test_version_and_loud_failure_without_gpu(kabc):
    L = kabc.lib()
    assert L.kabc_version() == 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kabc.KissABCError) as ei:   # no CPU fallback: creating a context must fail loudly
        kabc.Context()
    assert ei.value.code == 2


def test_pod_layouts_match_header(kabc):
    K = kabc._capi
    assert C.sizeof(K.PriorT) == 40 and C.sizeof(K.ModelT) == 16 + 32 * 8 + 8 * 8
    assert C.sizeof(K.SmcConfigT) == 72 and C.sizeof(K.AisConfigT) == 64 and C.sizeof(K.SmcLogT) == 56
    assert C.sizeof(K.AbcdeConfigT) == 48 and C.sizeof(K.PfilterConfigT) == 48
    from oracle import oracle as O
    for a, b in [(K.PriorT, O.Prior), (K.ModelT, O.Model), (K.SmcConfigT, O.SmcConfig), (K.AisConfigT, O.AisConfig), (K.SmcLogT, O.SmcLog),
                 (K.AbcdeConfigT, O.AbcdeConfig), (K.PfilterConfigT, O.PfilterConfig)]:
        assert [(n, t) for n, t in a._fields_] == [(n, t) for n, t in b._fields_]


def test_host_mirror_defaults_and_validation(kabc):
    cfg = kabc.smc_config()
    # ref src/smc.jl:95-105
    assert (cfg.nparticles, cfg.alpha, cfg.mcmc_retrys, cfg.mcmc_tol, cfg.epstol, cfg.max_stretch) == (100, 0.95, 0, 0.015, 0.0, 2.0)
    assert cfg.r_epstol == (1 - 0.95) ** 1.5 / 50 and cfg.min_r_ess == 0.95 * 0.95
    pri = kabc.Factored(kabc.Uniform(1, 3), kabc.Truncated(kabc.Normal(0, 0.1), 0, 100))
    assert len(pri) == 2
    pods = pri._pods()
    assert (pods[0].kind, pods[0].p0, pods[0].p1) == (0, 1.0, 3.0) and (pods[1].kind, pods[1].lo, pods[1].hi) == (2, 0.0, 100.0)
    # the laws and simulators added for the reference's integration tests (test/runtests.jl:46-56, :105-112)
    dp = kabc.Factored(kabc.NegativeBinomial(4.5, 0.13), kabc.Beta(15, 2), kabc.DiscreteUniform(1, 10))._pods()
    assert [(q.kind, q.p0, q.p1) for q in dp] == [(4, 4.5, 0.13), (3, 15.0, 2.0), (5, 1.0, 10.0)]
    sm, nm = kabc.Socks((0, 11), 11)._pod(), kabc.NoisyProduct(5.5, 0.01)._pod()
    assert (sm.kind, sm.n_target, sm.param[0]) == (5, 2, 11.0) and (nm.kind, nm.param[0], nm.param[1]) == (4, 2.0, 0.01)
    with pytest.raises(kabc.KissABCError):
        kabc.Factored()
    with pytest.raises(kabc.KissABCError):
        kabc.Truncated(kabc.Uniform(0, 1), 0, 1)
    with pytest.raises(kabc.KissABCError):
        kabc.smc(pri, lambda x: 0.0)          # arbitrary closures cannot run on the device
    with pytest.raises(kabc.KissABCError):
        kabc.GandK(target=(1, 2, 3))
    m = kabc.NormalMeanStd(1000, 2.0, 0.04, 50.0, precision="f64")._pod()
    assert (m.kind, m.precision, m.n_draws, m.target[0], m.target[1], m.param[0]) == (0, 0, 1000, 2.0, 0.04, 50.0)
    p = kabc.Particles([1.0, 2.0, 3.0])
    assert p.mean() == 2.0 and p.approx(2.5) and not p.approx(10.0)


def test_product_does_not_touch_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "kissabc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"(import|include|from|dlopen|CDLL)[^\n]*oracle|kor_[a-z]|libkabc_oracle", src), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libkissabc_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
    code = "import sys; sys.path.insert(0, %r); import kissabc_jl_b200; assert not any('oracle' in m for m in sys.modules)" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints one JSON line with the keys the driver reads (bounded sample, CPU only)."""
    import json
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--ref-budget", "1.0"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"] == "normal_smc"
