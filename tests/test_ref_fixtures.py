"""Reference-held golden vectors (tests/golden/ref_*.json, written by julia/make_ref_fixtures.jl from the UNMODIFIED KissABC.jl
running on julia/PhiloxRNG.jl) against the oracle's serial-stream mode.

No Julia exists in this repository's build image, so the fixtures are absent until a box with Julia produces them; every test
here then SKIPS (parity stays "unpinned at stream level").  When they are present:
  * integers must be identical: number of iterations, ESS after every cut, number of surviving particles -- any differing
    accept / resample decision changes them;
  * floats agree within RTOL = 1e-12 (north_star's FP64 bound).  They are not bit-identical by construction: the reference
    computes `log(rand(rng))` with Julia's libm log (<= 1 ulp from the spec's log), sums pairwise in mean/std where the spec
    sums sequentially (relative difference <= ~n eps/2 = 1.1e-14 at n = 100), and uses hypot where the spec uses sqrt(a^2+b^2)
    (<= 1 ulp): all three are bounded far below 1e-12 and are exactly the deviations VERDICT r1 lists.
The serial mode shares all control logic with the per-particle-stream mode the device is bit-compared with (tests/test_gpu_parity.py);
only the variate source differs (oracle/kabc_oracle.c, ST_SERIAL)."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.json")))
PYREF = sorted(glob.glob(os.path.join(HERE, "golden", "pyref_*.json")))
RTOL = 1e-12


def f64(bits):
    return np.array([int(b) for b in bits], dtype=np.uint64).view(np.float64)


def close(a, b, what):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, what
    fin = np.isfinite(a) & np.isfinite(b)
    assert (np.isfinite(a) == np.isfinite(b)).all() and (a[~fin] == b[~fin]).all(), what
    assert (np.abs(a[fin] - b[fin]) <= RTOL * np.maximum(np.abs(a[fin]), np.abs(b[fin])) + 1e-300).all(), what


def _objects(O, fx):
    pri = O.make_priors([tuple(p) for p in fx["prior"]])
    m = fx["model"]
    if m["kind"] == "normal":
        mod = O.make_model(O.NORMAL_MEANSTD, m["n"], (2.0, 0.04), (50.0,))
    elif m["kind"] == "noisyprod":  # test/runtests.jl:105-112: |(n*n + du) * (n + 0.01 randn) - 5.5|
        mod = O.make_model(O.DETERMINISTIC, 0, target=(5.5,), param=(2.0, 0.01))
    else:
        mod = O.make_model(O.MA2_AUTOCOV, m["n"], (0.72, 0.2))
    return pri, mod


def test_fixture_pipeline_is_in_place():
    """the generator and the RNG the first box with Julia needs are committed"""
    root = os.path.dirname(HERE)
    for f in ("julia/PhiloxRNG.jl", "julia/make_ref_fixtures.jl"):
        assert os.path.exists(os.path.join(root, f)), f


def test_serial_mode_is_a_pure_variate_source_change(oracle):
    """The serial mode consumes ONE stream in the reference's order; the control logic is shared: both modes run the same cut /
    resample / sweep code and only the variates differ."""
    O = oracle
    pri = O.make_priors([("uniform", -2, 2), ("uniform", -1, 1)])
    mod = O.make_model(O.MA2_AUTOCOV, 50, (0.72, 0.2))
    a = O.Smc(1234, pri, mod, O.smc_config(nparticles=200, alpha=0.9, max_iterations=6))
    b = O.Smc(1234, pri, mod, O.smc_config(nparticles=200, alpha=0.9, max_iterations=6))
    b.set_serial()
    a.init(); b.init()
    for _ in range(6):
        a.iterate(); b.iterate()
    la, lb = a.log(), b.log()
    assert len(la) == len(lb) == 6
    assert all(r["eps"] > 0 for r in lb) and lb[-1]["eps"] < lb[0]["eps"]
    assert all(0 < r["n_alive"] <= 200 for r in lb)
    tha, thb = a.state()[0], b.state()[0]
    assert not np.array_equal(tha, thb)  # different variates ...
    assert np.isfinite(thb).all()


@pytest.mark.parametrize("path", [p for p in FIXTURES if "ref_smc_" in p] or [None])
def test_reference_smc_run_matches_oracle_serial_mode(oracle, path):
    if path is None:
        pytest.skip("no tests/golden/ref_smc_*.json: run julia/make_ref_fixtures.jl on a box with Julia + KissABC 3.0.1")
    _check_smc_fixture(oracle, json.load(open(path)))


def _check_smc_fixture(O, fx):
    pri, mod = _objects(O, fx)
    kw = {}
    ints = set(fx["int_kwargs"])
    for k, v in fx["kwargs"].items():
        kw[k] = int(v) if k in ints else float(f64([v])[0])
    s = O.Smc(int(fx["seed"]), pri, mod, O.smc_config(**kw))
    s.set_serial()
    s.init()
    stop = 0
    while not stop:
        stop = s.iterate()
    log = s.log()
    assert len(log) == int(fx["iterations"])
    assert [r["n_alive"] for r in log] == list(fx["ess_per_iteration"])            # every cut kept the same particles
    close([r["eps"] for r in log], f64(fx["eps_per_iteration"]), "eps per iteration")
    th, X, lpi, alive = s.state()
    close(s.scalars()["eps"], f64([fx["eps"]])[0], "eps")
    close(X, f64(fx["C"]), "C = Xs of all particles (ref src/smc.jl:205)")
    for k, pk in enumerate(fx["P"]):
        row = th[k][alive.astype(bool)]
        if fx["prior"][k][0] == "duniform":  # the returned particles are push_p(prior, .), ref src/smc.jl:200
            row = np.rint(row)
        close(row, f64(pk), f"P[{k}] = theta of the alive particles")


@pytest.mark.parametrize("path", [p for p in FIXTURES if "ref_ais_" in p] or [None])
def test_reference_ais_transitions_match_oracle_serial_mode(oracle, path):
    if path is None:
        pytest.skip("no tests/golden/ref_ais_*.json: run julia/make_ref_fixtures.jl on a box with Julia + KissABC 3.0.1")
    _check_ais_fixture(oracle, json.load(open(path)))


def _check_ais_fixture(O, fx):
    pri, mod = _objects(O, fx)
    N, steps, nt = int(fx["nwalkers"]), int(fx["steps"]), int(fx["ntransitions"])
    scale = float(f64([fx["scale"]])[0])
    post = int(fx.get("posterior", 0))  # 1: ApproxPosterior (src/types.jl:76-104), `scale` = maxcost, the second slot holds the cost
    a = O.Ais(int(fx["seed"]), pri, mod, O.ais_config(N, steps + 1, ntransitions=nt, scale=scale, posterior=post))
    a.set_serial()
    a.init()
    th, lp, ll = a.state()
    close(th.ravel(), f64(fx["theta_init"]), "ensemble after the init step (ref src/KissABC.jl:50-61)")
    close(lp, f64(fx["lp_init"]), "log-prior after init"); close(ll, f64(fx["ll_init"]), "log-likelihood after init")
    a2 = O.Ais(int(fx["seed"]), pri, mod, O.ais_config(N, steps + 1, ntransitions=nt, scale=scale, posterior=post))
    a2.set_serial()
    out = a2.run_sequential()                                                     # the reference's own schedule, src/KissABC.jl:66-80
    ref = np.array([f64(s_) for s_ in fx["samples"]]).T                           # d x (steps+1)
    close(out, ref, "emitted samples")
    th, lp, ll = a2.state()
    close(th.ravel(), f64(fx["theta"]), "final ensemble"); close(lp, f64(fx["lp"]), "final log-prior"); close(ll, f64(fx["ll"]), "final ll")


# ---- the same comparisons against a SECOND restatement of the reference (tests/golden/make_pyref_fixtures.py): the Julia of
# src/smc.jl, src/transition.jl, src/types.jl:51-75, src/KissABC.jl:35-80 transliterated line by line into Python, with the
# platform's log / exp, pairwise mean / std and hypot like the reference.  Not the reference itself (parity stays unpinned), but
# every accept / cut / resample decision of five whole runs (1.3e7 Philox words) is taken identically by two independent codes.
@pytest.mark.parametrize("path", PYREF, ids=[os.path.basename(p)[:-5] for p in PYREF])
def test_python_transliteration_of_the_reference_matches_oracle_serial_mode(oracle, path):
    fx = json.load(open(path))
    (_check_smc_fixture if fx["kind"] == "smc" else _check_ais_fixture)(oracle, fx)


def test_python_transliteration_fixtures_are_committed():
    assert len(PYREF) == 7


def _b(x):
    return [str(int(v)) for v in np.atleast_1d(np.asarray(x, dtype=np.float64)).view(np.uint64)]


def test_loader_on_fixtures_in_the_julia_format(oracle):
    """Plumbing check of the loader (NOT a parity claim): fixtures written in the exact JSON layout of make_ref_fixtures.jl, but
    from the oracle's own serial mode, must pass the two comparisons above -- so that the first reference-made fixture is judged
    by a loader that is known to parse the format."""
    O = oracle
    seed = 0x4B49535341424300
    prior = [["uniform", -2, 2], ["uniform", -1, 1]]
    model = {"kind": "ma2", "n": 100}
    kw = dict(nparticles=300, alpha=0.9, epstol=0.3)
    fx = {"kind": "smc", "seed": str(seed), "prior": prior, "model": model,
          "kwargs": {"nparticles": "300", "alpha": _b(0.9)[0], "epstol": _b(0.3)[0]}, "int_kwargs": ["nparticles"]}
    pri, mod = _objects(O, fx)
    s = O.Smc(seed, pri, mod, O.smc_config(**kw))
    s.set_serial(); s.init()
    stop = 0
    while not stop:
        stop = s.iterate()
    th, X, lpi, alive = s.state()
    log = s.log()
    fx.update(iterations=str(len(log)), eps_per_iteration=_b([r["eps"] for r in log]), ess_per_iteration=[r["n_alive"] for r in log],
              eps=_b(s.scalars()["eps"])[0], C=_b(X), P=[_b(th[k][alive.astype(bool)]) for k in range(2)])
    _check_smc_fixture(O, json.loads(json.dumps(fx)))
    N, steps, nt, scale = 10, 25, 2, 0.2
    a = O.Ais(seed, pri, mod, O.ais_config(N, steps + 1, ntransitions=nt, scale=scale))
    a.set_serial(); a.init()
    th0, lp0, ll0 = a.state()
    a2 = O.Ais(seed, pri, mod, O.ais_config(N, steps + 1, ntransitions=nt, scale=scale))
    a2.set_serial()
    out = a2.run_sequential()
    th1, lp1, ll1 = a2.state()
    fa = {"kind": "ais", "seed": str(seed), "prior": prior, "model": model, "scale": _b(scale)[0], "nwalkers": str(N), "steps": str(steps),
          "ntransitions": str(nt), "theta_init": _b(th0.ravel()), "lp_init": _b(lp0), "ll_init": _b(ll0),
          "samples": [_b(out[:, m]) for m in range(steps + 1)], "theta": _b(th1.ravel()), "lp": _b(lp1), "ll": _b(ll1)}
    _check_ais_fixture(O, json.loads(json.dumps(fa)))
    assert (np.abs(out[:, 1:] - out[:, :-1]) > 0).any()  # the chain moved
