"""ABCDE / pfilter: the oracle against a literal Python transliteration of the reference.

The oracle (and the device, which is bit-compared with it) restates `ABCDE` and `pfilter` for a data-parallel machine: the
"uniform choice from a subset" of src/smc.jl:390 is taken over the (cost, index)-sorted population and pfilter's per-particle
`@goto resample` loop (src/smc.jl:308-322) runs as global rounds over the still-pending particles.  Both are re-orderings of
independent draws, so they must leave every distribution unchanged -- the reference holds no test of either sampler, so this
file checks exactly that: the reference's Julia, transliterated line by line (sequential loops, one numpy generator), and the
oracle are run on 24 seeds each on the README normal model, and every summary statistic of the result (posterior means and
spreads, cost levels, epsilon, the number of proposals / simulations) has to agree within Monte-Carlo error (|z| < 4; the
seeds are fixed, so the outcome is deterministic: the largest |z| today is 1.4)."""
import math

import numpy as np

AB = [(1.0, 3.0), (0.01, 0.2)]          # Factored(Uniform(1,3), Uniform(0.01,0.2))
NDRAW = 100
K = 24


def _logpdf(x):                          # src/priors.jl:30-37 + Distributions' Uniform
    s = 0.0
    for (a, b), v in zip(AB, x):
        s += -math.log(b - a) if a <= v <= b else -math.inf
    return s


def _prior_rand(rng):
    return np.array([a + (b - a) * rng.random() for a, b in AB])


def _cost_of(rng):                       # README.md:46-52
    def cost(th):
        x = rng.standard_normal(NDRAW) * th[1] + th[0]
        return float(np.hypot(x.mean() - 2.0, (x.std(ddof=1) - 0.04) * 50))
    return cost


def _init(rng, cost, N):                 # src/smc.jl:283-299 / :349-365 (identical in both samplers)
    th = [_prior_rand(rng) for _ in range(N)]
    lp = [_logpdf(th[i]) for i in range(N)]
    C = np.full(N, cost(th[0]))
    for i in range(N):
        if math.isfinite(lp[i]):
            C[i] = cost(th[i])
        while not math.isfinite(C[i]) or not math.isfinite(lp[i]):
            th[i] = _prior_rand(rng); lp[i] = _logpdf(th[i]); C[i] = cost(th[i])
    return th, lp, C


def ref_abcde(rng, eps_target, nparticles=50, generations=20, alpha=0.0, earlystop=False, proposal_width=1.0):
    """src/smc.jl:346-428, parallel = false"""
    cost, N = _cost_of(rng), nparticles
    th, lp, D = _init(rng, cost, N)
    nsims, iters = 0, 0
    g = proposal_width * 2.38 / math.sqrt(2 * len(AB))
    while iters < generations:
        iters += 1
        nth, nD, nlp = list(th), D.copy(), list(lp)
        el, eh = D.min(), D.max()
        if earlystop and eh <= eps_target:
            break
        epop = max(eps_target, el + alpha * (eh - el))
        for i in range(N):
            if earlystop and D[i] <= eps_target:
                continue
            s = i
            eps = eps_target if D[i] <= eps_target else epop
            if D[i] > eps:
                sub = np.nonzero(D <= D[i])[0]                    # (1:nparticles)[Δs .<= Δs[i]], index order
                s = int(sub[rng.integers(len(sub))])
            a = s
            while a == s:
                a = int(rng.integers(N))
            b = a
            while b == a or b == s:
                b = int(rng.integers(N))
            thp = th[s] + (th[a] - th[b]) * g
            l = _logpdf(thp)
            if math.log(rng.random()) > min(0, l - lp[i]):
                continue
            nsims += 1
            dp = cost(thp)
            if dp <= max(eps, D[i]):
                nD[i] = dp; nth[i] = thp; nlp[i] = l
        th, D, lp = nth, nD, nlp
    return np.array(th).T, D, nsims


def ref_pfilter(rng, N, q=0.7, eff_tol=0.1, epstol=-math.inf, max_iters=math.inf, proposal_width=0.75):
    """src/smc.jl:275-345, parallel = false"""
    cost = _cost_of(rng)
    lowN = 4 * len(AB)
    if N * q <= lowN:
        N = math.ceil((lowN + 1) / q)
    sm, lp, C = _init(rng, cost, N)
    iters, total = 0, 0
    while True:
        iters += 1
        eps = float(np.quantile(C, q))
        bad = C > eps
        idxok, idxbad = np.nonzero(~bad)[0], np.nonzero(bad)[0]
        nreps = 0
        for i in idxbad:
            while True:                                            # @label resample ... @goto resample
                b = c = d = int(idxok[rng.integers(len(idxok))])
                while c == b:
                    c = int(idxok[rng.integers(len(idxok))])
                while d == b or d == c:
                    d = int(idxok[rng.integers(len(idxok))])
                p = sm[b] + (sm[d] - sm[c]) * (rng.standard_normal() * proposal_width)
                nreps += 1
                ll = _logpdf(p)
                if math.log(rng.random()) > min(0.0, ll - lp[i]):
                    continue
                Cp = cost(p)
                if Cp > eps:
                    continue
                break
            C[i] = Cp; sm[i] = p; lp[i] = ll
        total += nreps
        if len(idxbad) / nreps < eff_tol or eps < epstol or iters > max_iters:
            break
    return np.array(sm).T, C, eps, iters, total


def _summary(th, C):
    return [th[0].mean(), th[1].mean(), th[0].std(), th[1].std(), C.mean(), C.max()]


def _z(a, b):
    a, b = np.array(a), np.array(b)
    se = np.sqrt(a.var(0) / len(a) + b.var(0) / len(b))
    return (a.mean(0) - b.mean(0)) / np.where(se > 0, se, 1.0)


def test_pfilter_is_distributed_like_the_reference_loop(oracle):
    O = oracle
    pri = O.make_priors([("uniform", 1, 3), ("uniform", 0.01, 0.2)])
    mod = O.make_model(O.NORMAL_MEANSTD, NDRAW, target=(2.0, 0.04), param=(50.0,))
    ref, orc = [], []
    for k in range(K):
        th, C, eps, iters, reps = ref_pfilter(np.random.default_rng(200 + k), 300, max_iters=12)
        ref.append(_summary(th, C) + [eps, iters, reps])
        r = O.pfilter(2000 + k, pri, mod, 300, max_iters=12)
        orc.append(_summary(r["theta"], r["C"]) + [r["eps"], r["iterations"], r["nreps"]])
    z = _z(ref, orc)
    assert (np.abs(z) < 4).all(), z
    assert np.array(ref)[:, 7].tolist() == np.array(orc)[:, 7].tolist()      # same number of iterations (max_iters + 1)


def test_abcde_is_distributed_like_the_reference_loop(oracle):
    O = oracle
    pri = O.make_priors([("uniform", 1, 3), ("uniform", 0.01, 0.2)])
    mod = O.make_model(O.NORMAL_MEANSTD, NDRAW, target=(2.0, 0.04), param=(50.0,))
    ref, orc = [], []
    for k in range(K):
        th, D, nsim = ref_abcde(np.random.default_rng(100 + k), 0.1, nparticles=200, generations=40, alpha=0.5)
        ref.append(_summary(th, D) + [nsim])
        r = O.abcde(1000 + k, pri, mod, 0.1, nparticles=200, generations=40, alpha=0.5)
        orc.append(_summary(r["theta"], r["C"]) + [r["nsim"]])
    z = _z(ref, orc)
    assert (np.abs(z) < 4).all(), z


def test_red_black_ais_has_the_stationary_law_of_the_reference_schedule(oracle):
    """Row a15's documented deviation: the device (and the oracle's `run_parallel`) moves the walkers in red/black half-steps where
    the reference moves one walker per `step` (src/KissABC.jl:66-80, the oracle's `run_sequential`).  Both are valid ensemble
    moves with partners from the frozen complementary set (src/transition.jl:51-59), so the stationary law is the same: 16
    seeds each on the README model, means and spreads of both parameters over the recorded chains agree within Monte-Carlo
    error (|z| < 4; fixed seeds, today's largest |z| is 0.3)."""
    O = oracle
    pri = O.make_priors([("uniform", 1, 3), ("uniform", 0.01, 0.2)])
    mod = O.make_model(O.NORMAL_MEANSTD, NDRAW, target=(2.0, 0.04), param=(50.0,))

    def stats(out):
        return [out[0].mean(), out[1].mean(), out[0].std(), out[1].std()]

    seq, par = [], []
    for k in range(16):
        a = O.Ais(500 + k, pri, mod, O.ais_config(16, 400, ntransitions=32, discard_initial=400, scale=0.05))
        seq.append(stats(a.run_sequential()))
        b = O.Ais(900 + k, pri, mod, O.ais_config(16, 400, ntransitions=8, discard_initial=400, scale=0.05))
        par.append(stats(b.run_parallel()))
    z = _z(seq, par)
    assert (np.abs(z) < 4).all(), z
    assert abs(np.mean(par, 0)[0] - 2.0) < 0.01 and abs(np.mean(par, 0)[1] - 0.0405) < 0.001
