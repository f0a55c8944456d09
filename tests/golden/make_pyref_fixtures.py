"""A second, structurally independent restatement of the reference's control flow, used to pin the C oracle's SERIAL mode.

    python tests/golden/make_pyref_fixtures.py        ->  tests/golden/pyref_{smc,ais}_*.json

This is julia/make_ref_fixtures.jl + julia/PhiloxRNG.jl + the reference's own `smc` / AIS `step` code transliterated LINE BY
LINE into Python (numpy semantics where the Julia code uses vector semantics: `quantile`, boolean masks, `repeat`), so that
the fixture pipeline of tests/test_ref_fixtures.py can run today.  It is NOT the reference (no Julia in the build image) and
makes no parity claim beyond: two independent restatements of src/smc.jl:119-205, src/transition.jl:1-82,
src/types.jl:51-75 and src/KissABC.jl:35-80 -- one in C written for speed and sharding, one in Python shaped like the Julia --
take identical decisions on identical variates.  Like the reference, and unlike the oracle, it uses the platform's `log`/`exp`,
numpy's pairwise `mean`/`std` and `hypot` (the three deviations tests/test_ref_fixtures.py bounds), so floats agree to 1e-12,
not bit for bit.  It shares with the oracle only the VARIATE SPEC primitives (Philox block, u01, Box-Muller pair, spec log for
`randexp`), called through oracle/libkabc_oracle.so; no control logic of the oracle is used.
The fixtures have exactly the JSON layout the Julia generator writes; files made by real Julia are named ref_*.json.
"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 0x4B49535341424300
ST_SERIAL = 6


# ------------------------------------------------------------------ julia/PhiloxRNG.jl
class PhiloxRNG:
    def __init__(self, seed, tag=ST_SERIAL, id_=0, epoch=0):
        self.L = O.lib()
        self.key = (C.c_uint32 * 2)(seed & 0xFFFFFFFF, seed >> 32)
        self.ctr = [0, id_, epoch, tag]
        self.buf, self.pos, self.consumed = None, 4, 0
        self.L.kor_u01.restype = C.c_double
        self.L.kor_log.restype = C.c_double

    def u32(self):
        if self.pos == 4:
            out = (C.c_uint32 * 4)()
            self.L.kor_philox4x32_10((C.c_uint32 * 4)(*self.ctr), self.key, out)
            self.buf = list(out)
            self.ctr[0] = (self.ctr[0] + 1) & 0xFFFFFFFF
            self.pos = 0
        self.pos += 1
        self.consumed += 1
        return self.buf[self.pos - 1]

    def u01(self, w):
        return (float(w) + 0.5) * 2.3283064365386962890625e-10

    def rand(self):                                  # rand(rng)
        return self.u01(self.u32())

    def rand_range(self, first, n):                  # rand(rng, first:first+n-1)
        return first + ((self.u32() * n) >> 32)

    def normal_pair(self, w0, w1):
        z0, z1 = C.c_double(), C.c_double()
        self.L.kor_normal_pair(C.c_uint32(w0), C.c_uint32(w1), C.byref(z0), C.byref(z1))
        return z0.value, z1.value

    def randn(self):                                 # randn(rng): two words, keeps r cos
        w0 = self.u32(); w1 = self.u32()
        return self.normal_pair(w0, w1)[0]

    def randexp(self):                               # randexp(rng) = -spec_log(u)
        return -self.L.kor_log(C.c_double(self.u01(self.u32())))

    def spec_normals(self, n):                       # 4 words -> 4 normals, leftovers of the last group dropped
        z = []
        while len(z) < n:
            w = [self.u32() for _ in range(4)]
            z.extend(self.normal_pair(w[0], w[1]))
            z.extend(self.normal_pair(w[2], w[3]))
        return np.array(z[:n])


# ------------------------------------------------------------------ cost closures of make_ref_fixtures.jl
def cost_normal(rng, n):
    def cost(theta):
        mu, sigma = theta
        x = rng.spec_normals(n) * sigma + mu
        return float(np.hypot(np.mean(x) - 2.0, (np.std(x, ddof=1) - 0.04) * 50))
    return cost


def cost_ma2(rng, n, target):
    def cost(theta):
        t1, t2 = theta
        if not (t1 > -2.0 and t1 < 2.0 and t1 + t2 > -1.0 and t1 - t2 < 1.0):
            return math.inf
        e = rng.spec_normals(n + 2)
        y = [e[t + 2] + t1 * e[t + 1] + t2 * e[t] for t in range(n)]
        a1 = 0.0
        for t in range(1, n):
            a1 += y[t] * y[t - 1]
        a2 = 0.0
        for t in range(2, n):
            a2 += y[t] * y[t - 2]
        return math.sqrt((a1 / n - target[0]) ** 2 + (a2 / n - target[1]) ** 2)
    return cost


# ------------------------------------------------------------------ Factored of Uniforms, src/priors.jl:30-43 + Distributions.jl
class FactoredUniform:
    def __init__(self, *ab):
        self.ab = ab

    def __len__(self):
        return len(self.ab)

    def rand(self, rng):                              # ntuple(i -> rand(rng, p[i])): a + (b - a) * rand(rng)
        return np.array([a + (b - a) * rng.rand() for a, b in self.ab])

    def logpdf(self, x):                              # s = logpdf(p[1], x[1]); s += ...
        s = None
        for (a, b), v in zip(self.ab, x):
            lp = -math.log(b - a) if a <= v <= b else -math.inf
            s = lp if s is None else s + lp
        return s


LOG2PI = 1.8378770664093453


class Factored:
    """Factored(Normal | Uniform | DiscreteUniform ...): rand / logpdf / push_p as Distributions.jl + src/priors.jl + src/types.jl:28-32"""
    def __init__(self, *comps):
        self.c = comps                                # ("normal", mu, sigma) | ("uniform", a, b) | ("duniform", a, b)

    def __len__(self):
        return len(self.c)

    def rand(self, rng):
        out = []
        for kind, p0, p1 in self.c:
            if kind == "normal":
                out.append(p0 + p1 * rng.randn())             # d.mu + d.sigma * randn(rng)
            elif kind == "uniform":
                out.append(p0 + (p1 - p0) * rng.rand())
            else:
                out.append(float(rng.rand_range(p0, p1 - p0 + 1)))   # rand(rng, d.a:d.b)
        return np.array(out)

    def push(self, x):                                # push_p: float(p) | round(Int, p) (ties to even, like Python's round)
        return np.array([float(round(v)) if c[0] == "duniform" else float(v) for c, v in zip(self.c, x)])

    def logpdf(self, x):
        s = None
        for (kind, p0, p1), v in zip(self.c, x):
            if kind == "normal":
                z = (v - p0) / p1
                lp = -(z * z + LOG2PI) / 2 - math.log(p1)     # StatsFuns.normlogpdf
            elif kind == "uniform":
                lp = -math.log(p1 - p0) if p0 <= v <= p1 else -math.inf
            else:
                pv = 1 / (p1 - p0 + 1)
                lp = math.log(pv) if (p0 <= v <= p1 and v == math.floor(v)) else -math.inf
            s = lp if s is None else s + lp
        return s


def cost_noisyprod(rng):                              # test/runtests.jl:105-112 with the run's rng instead of the global one
    def cost(theta):
        n, du = theta
        return abs((n * n + du) * (n + rng.randn() * 0.01) - 5.5)
    return cost


def quantile(v, p):
    """Statistics.quantile(v, p) (type 7: alpha = beta = 1), operation by operation -- numpy's `quantile` interpolates with a
    different rounding (b - (b-a)(1-t) for t >= 0.5), which is within 1e-12 but is not what the reference computes"""
    v = np.sort(v)
    n = len(v)
    m = 1.0 + p * (1.0 - 1.0 - 1.0)
    aleph = n * p + m
    j = min(max(int(aleph), 1), n - 1)
    g = min(max(aleph - j, 0.0), 1.0)
    a, b = (v[0], v[0]) if n == 1 else (v[j - 1], v[j])
    return float(a + g * (b - a)) if math.isfinite(a) and math.isfinite(b) else float((1 - g) * a + g * b)


# ------------------------------------------------------------------ src/smc.jl:92-206, transliterated (parallel = false)
def smc(prior, cost, rng, nparticles=100, alpha=0.95, mcmc_retrys=0, mcmc_tol=0.015, epstol=0.0, r_epstol=None,
        min_r_ess=None, max_stretch=2.0):
    r_epstol = (1 - alpha) ** 1.5 / 50 if r_epstol is None else r_epstol
    min_r_ess = alpha ** 2 if min_r_ess is None else min_r_ess
    Np = len(prior)
    N = nparticles
    push = getattr(prior, "push", lambda x: x)                                 # push_p(prior, .): identity for continuous laws
    th = [prior.rand(rng) for _ in range(N)]                                   # :119
    Xs = np.array([cost(push(th[i])) for i in range(N)])                       # :120-123
    lpis = np.array([prior.logpdf(push(th[i])) for i in range(N)])             # :125
    eps = math.inf
    alive = np.ones(N, dtype=bool)
    shown = []
    while True:
        epsv = eps
        eps = quantile(Xs[alive], alpha)                                       # :134
        flag = False
        if eps > Xs[alive].min():
            alive = Xs < eps
        else:
            alive = Xs <= eps
            flag = True
        ESS = int(alive.sum())
        shown.append((len(shown) + 1, eps, ESS))                               # verbose && @show iteration, eps, ESS
        if alpha * ESS <= N * min_r_ess:                                       # :145-153
            idxalive = np.nonzero(alive)[0]
            idx = np.tile(idxalive, math.ceil(N / len(idxalive)))[:N]          # repeat(idxalive, ceil(Int, N / n))[1:N]
            th = [th[j] for j in idx]
            Xs = Xs[idx]
            lpis = lpis[idx]
            alive = np.ones(N, dtype=bool)
        accepted = 0
        for _ in range(1 + mcmc_retrys):
            new_p = []
            for i in range(N):                                                 # :160-167, all proposals first
                if not alive[i]:
                    new_p.append(None)
                    continue
                a = b = i
                while a == i:
                    a = rng.rand_range(0, N)
                while b == i or b == a:
                    b = rng.rand_range(0, N)
                W = (th[b] - th[a]) * (max_stretch * rng.randn() / math.sqrt(Np))
                new_p.append((math.log(rng.rand()), th[i] + W, 0.0))
            for i in range(N):                                                 # :168-191
                if not alive[i]:
                    continue
                lprob, thp, logcorr = new_p[i]
                lpip = prior.logpdf(push(thp))
                if lpip < 0 and not math.isfinite(lpip):
                    continue
                d_ = lpip - lpis[i] + logcorr
                lM = d_ if d_ != d_ else min(d_, 0.0)                          # Julia's min propagates NaN
                if lprob < lM:
                    Xp = cost(push(thp))
                    if flag:
                        if Xp > eps:
                            continue
                    elif Xp >= eps:
                        continue
                    th[i] = thp
                    Xs[i] = Xp
                    lpis[i] = lpip
                    accepted += 1
            if accepted >= mcmc_tol * N:
                break
        if 2 * abs(epsv - eps) < r_epstol * (abs(epsv) + abs(eps)) or eps <= epstol or accepted < mcmc_tol * N:
            break
    P = np.array([push(th[i]) for i in range(N) if alive[i]]).T                # :200-204
    return dict(P=P, C=Xs, eps=eps, shown=shown)


# ------------------------------------------------------------------ AIS: src/transition.jl, src/types.jl:51-75, src/KissABC.jl:35-80
class KernelizedPosterior:
    def __init__(self, prior, cost, scale):
        self.prior, self.cost, self.scale = prior, cost, scale

    def loglike(self, x):                                                      # src/types.jl:51-58
        lp = self.prior.logpdf(x)
        ll = lp
        if math.isfinite(lp):
            ll = -0.5 * abs(self.cost(x) / self.scale) ** 2
        return (lp, ll)

    @staticmethod
    def valid(ld):
        return math.isfinite(ld[0] + ld[1])

    def accept(self, rng, old_ld, new_ld, corr):                               # src/types.jl:62-75
        if not math.isfinite(corr):
            raise RuntimeError("ld_correction is invalid")
        if not self.valid(old_ld):
            raise RuntimeError("starting sample invalid.")
        if not self.valid(new_ld):
            return False
        lW = corr + (new_ld[0] + new_ld[1]) - (old_ld[0] + old_ld[1])
        return -rng.randexp() <= lW


class HardPosterior:                                                           # ApproxPosterior, src/types.jl:76-104
    def __init__(self, prior, cost, maxcost):
        self.prior, self.cost, self.maxcost = prior, cost, maxcost

    def loglike(self, x):
        lp = self.prior.logpdf(x)
        cs = -lp
        if math.isfinite(lp):
            cs = self.cost(x)
        return (lp, cs)

    @staticmethod
    def valid(ld):
        return math.isfinite(ld[1]) and math.isfinite(ld[0])

    def accept(self, rng, old_ld, new_ld, corr):
        if not math.isfinite(corr):
            raise RuntimeError("ld_correction is invalid")
        if not self.valid(old_ld):
            raise RuntimeError("starting sample invalid.")
        if not self.valid(new_ld):
            return False
        lW = corr + new_ld[0] - old_ld[0]
        lW2 = max(self.maxcost, old_ld[1]) - new_ld[1]
        return (-rng.randexp() <= lW) and lW2 >= 0


def stretch_propose(rng, d, ps, i):                                            # src/transition.jl:50-59
    n = len(ps)
    a = i
    while i == a:
        a = rng.rand_range(0, n)
    u = rng.rand()
    Z = (u * (math.sqrt(3.0) - math.sqrt(1 / 3.0)) + math.sqrt(1 / 3.0)) ** 2  # cdf_g_inv(u, 3.0)
    W = (ps[i] - ps[a]) * Z
    return ps[a] + W, (d - 1) * math.log(Z)


def de_propose(rng, d, ps, i):                                                 # src/transition.jl:1-22
    n = len(ps)
    g = 2.38 / math.sqrt(2 * d) * math.exp(rng.randn() * 0.1)
    a = b = i
    while a == i:
        a = rng.rand_range(0, n)
    while b == a or b == i:
        b = rng.rand_range(0, n)
    W = (ps[a] - ps[b]) * g
    S = (np.abs(ps[a] - ps[b]) + np.abs(ps[i] - ps[b])) + np.abs(ps[a] - ps[i])
    T = np.array([g * x / 300 * rng.randn() for x in S])                       # one randn per component, in order
    return (ps[i] + W) + T, 0.0


def walk_propose(rng, d, ps, i):                                               # src/transition.jl:24-43
    n = len(ps)
    a = b = c = i
    while a == i:
        a = rng.rand_range(0, n)
    while b == a or b == i:
        b = rng.rand_range(0, n)
    while c == b or c == a or c == i:
        c = rng.rand_range(0, n)
    Xs = (ps[a] + (ps[b] + ps[c])) / 3
    z1 = rng.randn(); t1 = z1 * (ps[a] - Xs)
    z2 = rng.randn(); t2 = z2 * (ps[b] - Xs)
    z3 = rng.randn(); t3 = z3 * (ps[c] - Xs)
    return ps[i] + ((t1 + t2) + t3), 0.0


def transition(model, ps, lds, i, rng):                                        # src/transition.jl:61-82
    p = (1, 1, 1, 1, 2, 2, 3)[rng.rand_range(0, 7)]
    prop = (stretch_propose, de_propose, walk_propose)[p - 1]
    x, corr = prop(rng, len(model.prior), ps, i)
    ld = model.loglike(x)
    if model.accept(rng, lds[i], ld, corr):
        ps[i] = x
        lds[i] = ld
        return True
    return False


def ais_run(model, rng, N, steps, ntransitions, retry_sampling=100):
    ps = [model.prior.rand(rng) for _ in range(N)]                             # src/KissABC.jl:50
    lds = [model.loglike(ps[i]) for i in range(N)]                             # :51
    retrys = retry_sampling * N
    for i in range(N):                                                         # :53-61
        while not model.valid(lds[i]):
            ps[i] = model.prior.rand(rng)
            lds[i] = model.loglike(ps[i])
            retrys -= 1
            if retrys < 0:
                raise RuntimeError("Prior leads to ∞ costs too often, tune the prior or increase `retry_sampling`.")
    init = (np.array(ps).T.copy(), [ld[0] for ld in lds], [ld[1] for ld in lds])
    samples = [ps[N - 1].copy()]                                               # :63 push_p(model, particles[end])
    i = 0                                                                      # AISState(..., 1)
    for _ in range(steps):                                                     # :66-80
        for _ in range(ntransitions):
            transition(model, ps, lds, i, rng)
        samples.append(ps[i].copy())
        i = (i + 1) % N
    return init, samples, (np.array(ps).T.copy(), [ld[0] for ld in lds], [ld[1] for ld in lds])


# ------------------------------------------------------------------ fixtures in the layout of make_ref_fixtures.jl
def bits(x):
    return str(int(np.array([x], dtype=np.float64).view(np.uint64)[0]))


def jbits(v):
    return [bits(x) for x in np.asarray(v, dtype=np.float64).ravel()]


def smc_fixture(name, prior_spec, prior, mkcost, model_spec, **kw):
    rng = PhiloxRNG(SEED)
    res = smc(prior, mkcost(rng), rng, **kw)
    fx = {"kind": "smc", "name": name, "seed": str(SEED), "prior": prior_spec, "model": model_spec,
          "kwargs": {k: (str(v) if isinstance(v, int) else bits(v)) for k, v in kw.items()},
          "int_kwargs": [k for k, v in kw.items() if isinstance(v, int)],
          "iterations": str(len(res["shown"])), "eps_per_iteration": jbits([s[1] for s in res["shown"]]),
          "ess_per_iteration": [s[2] for s in res["shown"]], "eps": bits(res["eps"]), "C": jbits(res["C"]),
          "P": [jbits(res["P"][k]) for k in range(res["P"].shape[0])], "words_consumed": str(rng.consumed),
          "generator": "tests/golden/make_pyref_fixtures.py (Python transliteration of the reference; NOT Julia)"}
    json.dump(fx, open(os.path.join(HERE, f"pyref_smc_{name}.json"), "w"))
    print(f"pyref_smc_{name}: {len(res['shown'])} iterations, eps = {res['eps']}, alive = {res['P'].shape[1]}, words = {rng.consumed}")


def ais_fixture(name, prior_spec, prior, mkcost, model_spec, scale, N, steps, ntransitions, posterior=0):
    rng = PhiloxRNG(SEED)
    model = (HardPosterior if posterior else KernelizedPosterior)(prior, mkcost(rng), scale)
    (th0, lp0, ll0), samples, (th, lp, ll) = ais_run(model, rng, N, steps, ntransitions)
    fx = {"kind": "ais", "name": name, "seed": str(SEED), "prior": prior_spec, "model": model_spec, "scale": bits(scale),
          "nwalkers": str(N), "steps": str(steps), "ntransitions": str(ntransitions), "posterior": str(posterior),
          "theta_init": jbits(th0), "lp_init": jbits(lp0), "ll_init": jbits(ll0), "samples": [jbits(s) for s in samples],
          "theta": jbits(th), "lp": jbits(lp), "ll": jbits(ll), "words_consumed": str(rng.consumed),
          "generator": "tests/golden/make_pyref_fixtures.py (Python transliteration of the reference; NOT Julia)"}
    json.dump(fx, open(os.path.join(HERE, f"pyref_ais_{name}.json"), "w"))
    print(f"pyref_ais_{name}: {steps} steps x {ntransitions} transitions, words = {rng.consumed}")


UU_NORMAL = [["uniform", 1, 3], ["uniform", 0.01, 0.2]]
UU_MA2 = [["uniform", -2, 2], ["uniform", -1, 1]]
MA2_T = (0.72, 0.2)

if __name__ == "__main__":
    O.build()
    # the same seven cases as julia/make_ref_fixtures.jl
    smc_fixture("normal_defaults", UU_NORMAL, FactoredUniform((1, 3), (0.01, 0.2)), lambda r: cost_normal(r, 200),
                {"kind": "normal", "n": 200}, nparticles=400, epstol=0.05)
    smc_fixture("normal_sparse_resampling", UU_NORMAL, FactoredUniform((1, 3), (0.01, 0.2)), lambda r: cost_normal(r, 100),
                {"kind": "normal", "n": 100}, nparticles=300, alpha=0.8, min_r_ess=0.4, mcmc_retrys=2, mcmc_tol=0.3, epstol=0.1)
    smc_fixture("ma2", UU_MA2, FactoredUniform((-2, 2), (-1, 1)), lambda r: cost_ma2(r, 100, MA2_T),
                {"kind": "ma2", "n": 100}, nparticles=500, alpha=0.9, epstol=0.2)
    ais_fixture("normal", UU_NORMAL, FactoredUniform((1, 3), (0.01, 0.2)), lambda r: cost_normal(r, 100),
                {"kind": "normal", "n": 100}, 0.05, 12, 60, 3)
    ais_fixture("ma2", UU_MA2, FactoredUniform((-2, 2), (-1, 1)), lambda r: cost_ma2(r, 100, MA2_T),
                {"kind": "ma2", "n": 100}, 0.2, 10, 40, 2)
    smc_fixture("noisyprod_discrete", [["normal", 1, 0.5], ["duniform", 1, 10]], Factored(("normal", 1, 0.5), ("duniform", 1, 10)),
                cost_noisyprod, {"kind": "noisyprod"}, nparticles=200, epstol=0.02)
    ais_fixture("hard_normal", UU_NORMAL, FactoredUniform((1, 3), (0.01, 0.2)), lambda r: cost_normal(r, 100),
                {"kind": "normal", "n": 100}, 0.3, 12, 60, 3, posterior=1)       # ApproxPosterior(prior, cost, maxcost = 0.3)
