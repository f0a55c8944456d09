"""GPU vs the committed golden vectors (tests/golden/golden_v1.json): an anchor that does not need the oracle .so."""
import json
import os

import numpy as np
import pytest

from common import SEED

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")


def fh(s):
    return float.fromhex(s)


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_costs_match_golden(kabc, ctx, golden):
    W = kabc.workloads
    mk = {"normal": W.normal, "ma2": W.ma2, "gk": W.gk, "lv": W.lv, "socks": lambda prec, n: (None, kabc.Socks((0, 11), 11)),
          "noisyprod": lambda prec, n: (None, kabc.NoisyProduct(5.5, 0.01))}
    for key, v in golden["costs"].items():
        name, nd = key.rsplit("_", 1)
        _, cost = mk[name]("f64", int(nd))
        th = np.array([fh(x) for x in v["theta"]]).reshape(v["shape"])
        c = ctx.eval_cost(cost, th, first_id=5, epoch=9)
        assert [float(x).hex() for x in c] == v["cost"], key


def test_smc_run_matches_golden(kabc, ctx, golden):
    prior, cost = kabc.workloads.normal("f64", 100)
    s = kabc.SmcSession(ctx, prior, cost, kabc.smc_config(nparticles=256, max_iterations=12))
    s.init()
    while not s.iterate():
        pass
    g = golden["smc_normal_256"]
    log = s.log()
    assert [float(r["eps"]).hex() for r in log] == g["eps"]
    assert [r["n_alive"] for r in log] == g["n_alive"] and [r["accepted"] for r in log] == g["accepted"]
    th, X, _, _ = s.state()
    assert [float(x).hex() for x in th[:, :8].ravel()] == g["theta_first8"]
    assert [float(x).hex() for x in X[:8]] == g["X_first8"]
    assert float(th.sum()).hex() == g["theta_sum"] and s.scalars()["cost_evals"] == g["cost_evals"]


def test_ais_run_matches_golden(kabc, ctx, golden):
    prior, cost = kabc.workloads.normal("f64", 100)
    post = kabc.ApproxKernelizedPosterior(prior, cost, 0.05)
    res, cnt = kabc.sample(post, kabc.AIS(12), 40, ntransitions=5, discard_initial=3, thinning=2, ctx=ctx, return_counters=True)
    out = np.vstack([p.particles for p in res])
    assert [float(x).hex() for x in out.ravel()] == golden["ais_normal_12"]["samples"]
    assert cnt["cost_evals"] == golden["ais_normal_12"]["counters"]["cost_evals"]
    assert cnt["accepted"] == golden["ais_normal_12"]["counters"]["accepted"]


def test_added_laws_and_socks_match_golden(kabc, ctx, golden):
    """Beta / NegativeBinomial / DiscreteUniform draws and densities, and the reference's socks test at reduced size."""
    g = golden["priors_ext"]
    mk = {"beta": kabc.Beta, "negbin": kabc.NegativeBinomial, "duniform": kabc.DiscreteUniform}
    laws = [mk[s_[0]](*s_[1:]) for s_ in g["spec"]]
    n = g["shape"][1]
    th = ctx.prior_sample(kabc.Factored(*laws), n, first_id=0, epoch=7)
    assert [float(x).hex() for x in th.ravel()] == g["draws"]
    lps = []
    for k, law in enumerate(laws):
        xs = np.array(list(th[k]) + [0.0, 1.0, 2.5, -1.0, 7.0])
        lps += list(ctx.prior_logpdf(law, xs[None, :]))
    assert [float(x).hex() for x in lps] == g["logpdf"]
    R = -30.0 ** 2 / (30.0 - 15.0 ** 2)
    pri = kabc.Factored(kabc.NegativeBinomial(R, R / (30.0 + R)), kabc.Beta(15, 2))
    s = kabc.SmcSession(ctx, pri, kabc.Socks((0, 11), 11), kabc.smc_config(nparticles=500, alpha=0.99, r_epstol=0, epstol=0.01))
    s.init()
    while not s.iterate():
        pass
    gs = golden["smc_socks_500"]
    log = s.log()
    assert [float(r["eps"]).hex() for r in log] == gs["eps"]
    assert [r["n_alive"] for r in log] == gs["n_alive"] and [r["accepted"] for r in log] == gs["accepted"]
    th, X, _, _ = s.state()
    assert float(th.sum()).hex() == gs["theta_sum"] and float(X.sum()).hex() == gs["X_sum"]
    assert s.scalars()["cost_evals"] == gs["cost_evals"]
    res, cnt = kabc.sample(kabc.ApproxPosterior(pri, kabc.Socks((0, 11), 11), 0.1), kabc.AIS(50), 200, ntransitions=10, ctx=ctx,
                           return_counters=True)
    out = np.vstack([p.particles for p in res])
    assert [float(x).hex() for x in out.ravel()] == golden["ais_socks_50"]["samples"]
    assert cnt["cost_evals"] == golden["ais_socks_50"]["counters"]["cost_evals"]


def test_abcde_pfilter_match_golden(kabc, ctx, golden):
    prior, cost = kabc.workloads.normal("f64", 100)
    r = kabc.ABCDE(prior, cost, 0.05, nparticles=300, generations=25, alpha=0.3, ctx=ctx)
    g = golden["abcde_normal_300"]
    assert [float(x).hex() for x in np.vstack([p.particles for p in r.P]).ravel()] == g["theta"]
    assert [float(x).hex() for x in r.C.particles] == g["C"] and r.nsim == g["nsim"] and r.reached_eps == g["reached"]
    R = -30.0 ** 2 / (30.0 - 15.0 ** 2)
    pri = kabc.Factored(kabc.NegativeBinomial(R, R / (30.0 + R)), kabc.Beta(15, 2))
    f = kabc.pfilter(pri, kabc.Socks((0, 11), 11), 400, max_iters=4, ctx=ctx)
    g = golden["pfilter_socks_400"]
    assert [float(x).hex() for x in np.vstack([p.particles for p in f.P]).ravel()] == g["theta"]
    assert [float(x).hex() for x in f.C.particles] == g["C"] and float(f.eps).hex() == g["eps"]
    assert (f.nreps, f.iterations, f.cost_evals) == (g["nreps"], g["iterations"], g["cost_evals"])


def test_select_handles_ties_and_infinities(kabc, ctx):
    """bucket select edge cases: all-equal costs (slow path), many +Inf, tiny populations."""
    # deterministic cost |theta - 1.5| with a 2-point-like prior -> massive ties after resampling
    post_prior = kabc.Uniform(1.4999999, 1.5000001)
    res = kabc.smc(post_prior, kabc.Deterministic(1, 1.5), nparticles=5000, alpha=0.5, max_iterations=30, ctx=ctx)
    assert np.isfinite(res.eps) and res.eps >= 0 and len(res.P) > 0
    # smallest legal population
    res = kabc.smc(kabc.Normal(1, 0.2), kabc.Deterministic(0, 1.5), nparticles=4, max_iterations=5, ctx=ctx)
    assert res.C.shape == (4,)
