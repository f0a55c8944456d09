"""Full-size (BASELINE.json sizes) GPU checks through size-independent properties: the oracle cannot run 2^20 particles
x 1000 draws in seconds, so each smc iteration is verified against its own definition (src/smc.jl:131-191) with numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N = 1 << 20


def _type7(v_sorted, p):
    """Statistics.quantile (type 7) in the exact operation order of the oracle / device"""
    n = len(v_sorted)
    aleph = n * p + (1.0 - p)
    j = min(max(int(aleph), 1), n - 1)
    g = min(max(aleph - j, 0.0), 1.0)
    a, b = (v_sorted[0], v_sorted[0]) if n == 1 else (v_sorted[j - 1], v_sorted[j])
    return a + g * (b - a) if np.isfinite(a) and np.isfinite(b) else (1.0 - g) * a + g * b


# lv_smc = BASELINE.json config 5 at its full size (Lotka-Volterra SSA, 3 rates, 2^20 particles)
@pytest.mark.parametrize("wl,cfg", [("normal_smc", dict()), ("ma2_smc", dict(alpha=0.7, min_r_ess=0.3)), ("lv_smc", dict())])
def test_smc_iteration_properties_at_2_pow_20(kabc, ctx, wl, cfg):
    prior, cost = kabc.workloads.WORKLOADS[wl]("f32")
    d = len(prior)
    alpha = cfg.get("alpha", 0.95)
    s = kabc.SmcSession(ctx, prior, cost, kabc.smc_config(nparticles=N, **cfg))
    s.trace_enable(True)
    s.init()
    th, X, lpi, alive = s.state()
    assert alive.all() and th.shape == (d, N)
    lp_dev = ctx.prior_logpdf(prior, th)
    assert (lp_dev.view(np.uint64) == lpi.view(np.uint64)).all()            # lpi = logpdf(prior, theta), ref :125
    evals = N
    for it in range(4):
        th0, X0, lpi0, alive0 = th, X, lpi, alive
        s.iterate()
        th, X, lpi, alive = s.state()
        sc, tr, log = s.scalars(), s.trace(), s.log()[-1]
        # ref :134-142 -- epsilon is the exact type-7 quantile of the alive costs, bit for bit
        eps = _type7(np.sort(X0[alive0 == 1]), alpha)
        assert np.float64(eps).view(np.uint64) == np.float64(sc["eps"]).view(np.uint64)
        flag = 0 if eps > X0[alive0 == 1].min() else 1
        assert sc["flag"] == flag
        cut = (X0 <= eps) if flag else (X0 < eps)
        assert log["n_alive"] == int(cut.sum())
        # ref :145-153 -- resample decision and the cyclic tiling
        resample = alpha * float(cut.sum()) <= N * cfg.get("min_r_ess", alpha * alpha)
        assert bool(log["resampled"]) == resample
        idx = np.nonzero(cut)[0][np.arange(N) % int(cut.sum())] if resample else np.arange(N)
        alive_pre = np.ones(N, bool) if resample else cut
        # ref :160-191 -- per-particle decisions of the sweep, from the device's own trace
        dec = tr["decision"]
        assert ((dec == 0) == ~alive_pre).all()
        acc = dec == 4
        assert log["accepted"] == int(acc.sum())
        evals += int(((dec == 3) | (dec == 4)).sum())
        assert sc["cost_evals"] == evals
        a, b, z = tr["a"], tr["b"], tr["z"]
        live = alive_pre
        ii = np.arange(N)
        assert ((a[live] != ii[live]) & (b[live] != ii[live]) & (a[live] != b[live])).all()
        assert (a[live] >= 0).all() and (a[live] < N).all() and (b[live] < N).all()
        # proposal arithmetic: theta_i + (theta_b - theta_a) * (max_stretch*z/sqrt(Np)), same operation order
        src = th0[:, idx]
        scl = (2.0 * z[live]) / np.sqrt(float(d))
        thp = src[:, live] + (src[:, b[live]] - src[:, a[live]]) * scl
        assert (thp.view(np.uint64) == tr["theta_p"][:, live].view(np.uint64)).all()
        # accept rule: Xp < eps (<= when flag); rows of accepted particles are the proposals, the others are the gathered rows
        xp = tr["xp"]
        sim = (dec == 3) | (dec == 4)
        assert ((xp[sim] <= eps) if flag else (xp[sim] < eps)).tolist() == acc[sim].tolist()
        expect_th = np.where(acc, tr["theta_p"], src)
        expect_X = np.where(acc, xp, X0[idx])
        assert (expect_th.view(np.uint64) == th.view(np.uint64)).all()
        assert (expect_X.view(np.uint64) == X.view(np.uint64)).all()
        assert (np.where(acc, tr["lpi_p"], lpi0[idx]).view(np.uint64) == lpi.view(np.uint64)).all()
        assert (alive == alive_pre).all()
        # decision 1 <=> proposal outside the prior support (lpi_p = -Inf); prior-MH pre-test otherwise (ref :172-175)
        assert (np.isneginf(tr["lpi_p"][dec == 1])).all() and np.isfinite(tr["lpi_p"][sim]).all()
        lM = np.minimum(tr["lpi_p"] - lpi0[idx], 0.0)
        assert (tr["lprob"][sim] < lM[sim]).all() and (tr["lprob"][dec == 2] >= lM[dec == 2]).all()
    assert s.scalars()["iteration"] == 4


def test_smc_whole_run_idempotent_and_monotone(kabc, ctx):
    """same seed -> bit-identical run; epsilon decreases monotonically; the result obeys C[alive] < eps (ref :200-205)"""
    prior, cost = kabc.workloads.normal("f32")
    r1 = kabc.smc(prior, cost, nparticles=N, epstol=0.02, ctx=ctx)
    r2 = kabc.smc(prior, cost, nparticles=N, epstol=0.02, ctx=ctx)
    assert r1.eps == r2.eps and r1.iterations == r2.iterations and r1.cost_evals == r2.cost_evals
    assert (r1.C.view(np.uint64) == r2.C.view(np.uint64)).all()
    assert (r1.P[0].particles.view(np.uint64) == r2.P[0].particles.view(np.uint64)).all()
    eps = [r["eps"] for r in r1.log]
    assert all(x > y for x, y in zip(eps, eps[1:])) and eps[-1] <= 0.02
    assert len(r1.P[0]) == N and (r1.C < eps[-2]).all()
    assert abs(r1.P[0].mean() - 2.0) < 1e-3 and abs(r1.P[1].mean() - 0.04) < 2e-4


def test_ais_fullsize_sweep_properties(kabc, ctx):
    """2^18 walkers (config 4 size) of the normal model: red/black partners, accept rule and move mixture"""
    Nw = 1 << 18
    prior, cost = kabc.workloads.normal("f32")
    a = kabc.AisSession(ctx, prior, cost, kabc.ais_config(Nw, 1, scale=0.05))
    a.trace_enable(True)
    a.init()
    th0, lp0, ll0 = a.state()
    assert np.isfinite(lp0 + ll0).all()
    a.sweep(1)
    th, lp, ll = a.state()
    t = a.trace()
    h = Nw // 2
    i = np.arange(Nw)
    comp_lo = np.where(i < h, h, 0)
    for key, need in (("a", t["move"] >= 1), ("b", t["move"] >= 2), ("c", t["move"] == 3)):
        p = t[key]
        assert ((p[need] >= comp_lo[need]) & (p[need] < comp_lo[need] + h)).all()   # complementary colour only
        assert (p[~need] == -1).all()
    frac = np.bincount(t["move"], minlength=4)[1:] / Nw
    assert np.allclose(frac, [4 / 7, 2 / 7, 1 / 7], atol=0.005)                        # ref src/transition.jl:62
    acc = t["decision"] == 2
    moved = (th != th0).any(axis=0)
    assert (moved <= acc).all() and acc.sum() > 0
    # accept rule, ref src/types.jl:74 on the recorded variates; the colour-1 walkers saw colour 0 already moved
    first = i < h
    lW = (t["corr"] + (t["lp_p"] + t["ll_p"])) - (lp0 + ll0)
    valid = t["decision"] > 0
    assert ((-t["e"][valid & first] <= lW[valid & first]) == acc[valid & first]).all()
    Z = np.exp(t["corr"][t["move"] == 1])
    assert (Z > 1 / 3 - 1e-9).all() and (Z < 3 + 1e-9).all()


def test_gk_ais_fullsize_sweep_with_oracle_slice(oracle, kabc, ctx):
    """BASELINE.json config 4 at its full size: g-and-k (4 parameters, 10^4 draws, octile distance), AIS with 2^18 walkers, one
    red/black sweep.  Invariants of the whole ensemble from the device's own trace, plus 256 walkers re-evaluated by the
    oracle: cost of the recorded proposal on the walker's own Philox stream (F32 tolerance) and the accept rule."""
    import ctypes as C
    from common import SEED, models
    from test_gpu_parity import F32_TOL
    O = oracle
    Nw = 1 << 18
    prior, cost = kabc.workloads.gk("f32")
    scale = 0.5
    a = kabc.AisSession(ctx, prior, cost, kabc.ais_config(Nw, 1, scale=scale))
    a.trace_enable(True)
    a.init()
    th0, lp0, ll0 = a.state()
    assert th0.shape == (4, Nw) and np.isfinite(lp0 + ll0).all()
    a.sweep(1)
    th, lp, ll = a.state()
    t = a.trace()
    cn = a.counters()
    h = Nw // 2
    i = np.arange(Nw)
    comp_lo = np.where(i < h, h, 0)
    for key, need in (("a", t["move"] >= 1), ("b", t["move"] >= 2), ("c", t["move"] == 3)):
        p = t[key]
        assert ((p[need] >= comp_lo[need]) & (p[need] < comp_lo[need] + h)).all()
    frac = np.bincount(t["move"], minlength=4)[1:] / Nw
    assert np.allclose(frac, [4 / 7, 2 / 7, 1 / 7], atol=0.005)
    acc = t["decision"] == 2
    valid = t["decision"] > 0
    assert cn["cost_evals"] == Nw + int(np.isfinite(t["lp_p"]).sum())          # one evaluation per in-support proposal
    assert cn["accepted"] == int(acc.sum()) and acc.sum() > 0
    # accepted walkers hold their proposal, the others did not move
    assert (np.where(acc, t["theta_p"], th0).view(np.uint64) == th.view(np.uint64)).all()
    first = i < h
    lW = (t["corr"] + (t["lp_p"] + t["ll_p"])) - (lp0 + ll0)
    assert ((-t["e"][valid & first] <= lW[valid & first]) == acc[valid & first]).all()
    # oracle slice: 256 walkers of the first colour whose proposal reached the simulator
    M = models(O, kabc)["gk"]
    m = M["omodel"]()
    pick = np.nonzero(valid & first)[0][:: max(1, int((valid & first).sum()) // 256)][:256]
    atol, rtol = F32_TOL["gk"]
    for w in pick:
        tp = np.ascontiguousarray(t["theta_p"][:, w])
        c = O.lib().kor_cost(C.byref(m), SEED, 4, tp.ctypes.data_as(C.POINTER(C.c_double)), int(w), 0)  # epoch 0 = first half-step
        ll_ref = -0.5 * (c / scale) ** 2
        ll_dev = t["ll_p"][w]
        c_dev = scale * np.sqrt(-2.0 * ll_dev)
        assert abs(c_dev - c) <= atol + rtol * abs(c) + 1e-9, (w, c_dev, c)
        assert np.isfinite(ll_ref)
