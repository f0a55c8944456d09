"""CPU tests (no GPU): the oracle against the reference's own exact values, published algorithms (numpy), the
committed golden vectors and the statistical fixtures of the reference's tests and README."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from common import SEED, models, prior_draws

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")


def fh(s):
    return float.fromhex(s)


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


# ------------------------------------------------------------------ variate spec
def py_philox(ctr, key):
    """independent pure-Python Philox4x32-10 (Salmon et al. 2011)"""
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xffffffff, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xffffffff]
        k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
    return tuple(c)


def test_philox_known_answers(oracle, golden):
    for v in golden["philox_kat"]:
        assert oracle.philox(v["ctr"], v["key"]) == tuple(v["out"])
        assert py_philox(v["ctr"], v["key"]) == tuple(v["out"])
    rng = np.random.default_rng(1)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, 2**32, 4)]
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert oracle.philox(ctr, key) == py_philox(ctr, key)


def test_stream_layout(oracle, golden):
    """stream (seed; block j, id, epoch, tag): word k is word k%4 of Philox(ctr=(k//4, id, epoch, tag), key=seed)."""
    L = oracle.lib()
    key = (SEED & 0xffffffff, SEED >> 32)
    for tag, pid, ep, kk, w in golden["stream_words"]:
        assert L.kor_stream_word(SEED, tag, pid, ep, kk) == w
        assert py_philox((kk // 4, pid, ep, tag), key)[kk % 4] == w


def test_elementary_functions_accuracy_and_golden(oracle, golden):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.uniform(2.0**-33, 1, 5000), rng.uniform(1, 1e6, 2000), [1 / 3, 3.0]]):
        assert abs(L.kor_log(float(x)) - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x)))
    assert L.kor_log(1.0) == 0.0 and L.kor_log(0.0) == -math.inf and math.isnan(L.kor_log(-1.0))
    for x in rng.uniform(-60, 60, 5000):
        assert abs(L.kor_exp(float(x)) / math.exp(x) - 1) < 5e-16
    assert L.kor_exp(0.0) == 1.0 and L.kor_exp(800.0) == math.inf and L.kor_exp(-800.0) == 0.0
    s, c = C.c_double(), C.c_double()
    for w in rng.integers(0, 2**32, 5000):
        u = (float(w) + 0.5) * 2.0**-32
        L.kor_sincos2pi(u, C.byref(s), C.byref(c))
        assert abs(s.value - math.sin(2 * math.pi * u)) < 1.5e-15 and abs(c.value - math.cos(2 * math.pi * u)) < 1.5e-15
        assert abs(s.value**2 + c.value**2 - 1) < 1e-15
    for x, y in golden["log"]:
        assert L.kor_log(fh(x)) == fh(y)
    for x, y in golden["exp"]:
        assert L.kor_exp(fh(x)) == fh(y)


def test_uniform_index_normal(oracle, golden):
    L = oracle.lib()
    assert L.kor_u01(0) == 0.5 * 2.0**-32 and L.kor_u01(2**32 - 1) == (2**32 - 0.5) * 2.0**-32 < 1.0
    assert L.kor_index(0, 10) == 0 and L.kor_index(2**32 - 1, 10) == 9 and L.kor_index(2**31, 7) == 3
    z0, z1 = C.c_double(), C.c_double()
    for w0, w1, a, b in golden["normal_pairs"]:
        L.kor_normal_pair(w0, w1, C.byref(z0), C.byref(z1))
        assert z0.value == fh(a) and z1.value == fh(b)
    rng = np.random.default_rng(3)
    zs = []
    for w0, w1 in rng.integers(0, 2**32, (20000, 2)):
        L.kor_normal_pair(int(w0), int(w1), C.byref(z0), C.byref(z1))
        u1, u2 = (w0 + 0.5) * 2.0**-32, (w1 + 0.5) * 2.0**-32
        r = math.sqrt(-2 * math.log(u1))
        assert abs(z0.value - r * math.cos(2 * math.pi * u2)) < 1e-14 and abs(z1.value - r * math.sin(2 * math.pi * u2)) < 1e-14
        zs += [z0.value, z1.value]
    zs = np.array(zs)
    assert abs(zs.mean()) < 0.02 and abs(zs.std() - 1) < 0.02 and abs(((zs - zs.mean())**4).mean() / zs.var()**2 - 3) < 0.1


# ------------------------------------------------------------------ priors (ref test/runtests.jl:8-22)
def lp(oracle, pri, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return oracle.lib().kor_prior_logpdf(pri, len(x), x.ctypes.data_as(C.POINTER(C.c_double)))


def test_factored_exact_values_of_the_reference(oracle):
    d = oracle.make_priors([("uniform", 0, 1), ("uniform", 100, 101)])
    assert math.exp(lp(oracle, d, [0.0, 0.0])) == 0.0          # pdf(d,(0.0,0.0)) == 0.0
    assert math.exp(lp(oracle, d, [0.5, 100.5])) == 1.0        # pdf(d,(0.5,100.5)) == 1.0
    assert lp(oracle, d, [0.5, 100.5]) == 0.0                  # logpdf == 0.0
    assert lp(oracle, d, [0.0, 0.0]) == -math.inf              # logpdf == -Inf
    buf = (C.c_double * 2)()
    for i in range(200):
        assert oracle.lib().kor_prior_sample(SEED, d, 2, i, 0, buf) == 0
        assert 0 < buf[0] < 1 and 100 < buf[1] < 101


def test_prior_logpdf_against_scipy(oracle):
    from scipy import stats
    pri = oracle.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100), ("normal", -1, 2.5)])
    rng = np.random.default_rng(0)
    for _ in range(500):
        x = [rng.uniform(1, 3), abs(rng.normal(0, 0.1)), rng.normal(-1, 5)]
        ref = stats.uniform(1, 2).logpdf(x[0]) + stats.truncnorm(0, 1000, 0, 0.1).logpdf(x[1]) + stats.norm(-1, 2.5).logpdf(x[2])
        assert abs(lp(oracle, pri, x) - ref) < 1e-12 * max(1, abs(ref))
    assert lp(oracle, pri, [2, -0.01, 0]) == -math.inf and lp(oracle, pri, [3.0001, 0.1, 0]) == -math.inf


def test_prior_sampling_moments(oracle):
    th = prior_draws(oracle, [("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100), ("normal", -1, 2.5)], 20000)
    assert abs(th[0].mean() - 2) < 0.02 and th[0].min() > 1 and th[0].max() < 3
    assert th[1].min() >= 0 and abs(th[1].mean() - 0.1 * math.sqrt(2 / math.pi)) < 0.002  # half-normal mean
    assert abs(th[2].mean() + 1) < 0.06 and abs(th[2].std() - 2.5) < 0.05


def test_lgamma_and_added_laws_against_scipy(oracle, golden):
    """Beta / NegativeBinomial / DiscreteUniform (ref test/runtests.jl:50-51,106) + the spec'd log-gamma they use."""
    from scipy import special, stats
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for x in np.concatenate([rng.uniform(1e-6, 30, 5000), rng.uniform(30, 1e5, 1000), [1.0, 2.0, 0.5, 10.0, 9.999999]]):
        r = special.gammaln(x)
        assert abs(L.kor_lgamma(float(x)) - r) <= 1e-14 * max(1.0, abs(r))
    for x, y in golden["lgamma"]:
        assert L.kor_lgamma(fh(x)) == fh(y)
    from common import SOCKS_P, SOCKS_R
    cases = [(("beta", 15, 2), stats.beta(15, 2).logpdf), (("beta", 0.5, 0.7), stats.beta(0.5, 0.7).logpdf),
             (("beta", 1, 3), stats.beta(1, 3).logpdf), (("negbin", SOCKS_R, SOCKS_P), stats.nbinom(SOCKS_R, SOCKS_P).logpmf),
             (("negbin", 0.6, 0.9), stats.nbinom(0.6, 0.9).logpmf), (("duniform", 1, 10), stats.randint(1, 11).logpmf)]
    for spec, ref in cases:
        pri = oracle.make_priors([spec])
        xs = rng.uniform(0.001, 0.999, 300) if spec[0] == "beta" else np.arange(0, 300, dtype=float)
        for x in xs:
            a, b = lp(oracle, pri, [x]), ref(x)
            assert (a == b == -math.inf) or abs(a - b) < 1e-11 * max(1.0, abs(b)), (spec, x, a, b)
    # outside the support / non-integer arguments of the discrete laws
    assert lp(oracle, oracle.make_priors([("beta", 15, 2)]), [1.5]) == -math.inf
    assert lp(oracle, oracle.make_priors([("negbin", 3, 0.5)]), [-1.0]) == -math.inf
    assert lp(oracle, oracle.make_priors([("negbin", 3, 0.5)]), [2.5]) == -math.inf
    assert lp(oracle, oracle.make_priors([("duniform", 1, 10)]), [3.5]) == -math.inf
    assert lp(oracle, oracle.make_priors([("duniform", 1, 10)]), [11.0]) == -math.inf
    assert abs(lp(oracle, oracle.make_priors([("beta", 1, 3)]), [0.0]) - math.log(3.0)) < 1e-15   # xlogy(0, 0) = 0


def test_added_laws_sampling_distributions(oracle):
    from scipy import stats
    from common import SOCKS_P, SOCKS_R
    n = 20000
    th = prior_draws(oracle, [("negbin", SOCKS_R, SOCKS_P), ("beta", 15, 2), ("duniform", 1, 10), ("beta", 0.5, 0.7),
                              ("negbin", 0.6, 0.9)], n)
    assert stats.kstest(th[1], stats.beta(15, 2).cdf).pvalue > 1e-3 and stats.kstest(th[3], stats.beta(0.5, 0.7).cdf).pvalue > 1e-3
    for row, law in [(th[0], stats.nbinom(SOCKS_R, SOCKS_P)), (th[4], stats.nbinom(0.6, 0.9)), (th[2], stats.randint(1, 11))]:
        assert (row == np.rint(row)).all() and row.min() >= 0
        k, c = np.unique(row, return_counts=True)
        e = law.pmf(k) * n
        sel = e > 20
        chi2 = ((c[sel] - e[sel]) ** 2 / e[sel]).sum()
        assert chi2 < sel.sum() + 5 * math.sqrt(2 * sel.sum()) + 5, (chi2, sel.sum())
    assert abs(th[0].mean() - 30) < 0.5 and abs(th[0].std() - 15) < 0.5       # prior_mu, prior_sd of the reference test


def test_push_p(oracle):
    """ref test/runtests.jl:24-31 and src/types.jl:28-32: continuous components pass, discrete ones round (ties to even)."""
    L = oracle.lib()
    pri = oracle.make_priors([("normal", 0, 1), ("duniform", 0, 1), ("negbin", 3, 0.5), ("beta", 2, 2)])
    x = np.array([2.0, 1.0, 2.5, 0.3])
    out = np.empty(4)
    L.kor_push_p(pri, 4, x.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert list(out) == [2.0, 1.0, 2.0, 0.3]
    x = np.array([-0.49, 3.5, -0.5, 0.3])
    L.kor_push_p(pri, 4, x.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert list(out) == [-0.49, 4.0, -0.0, 0.3]


def test_added_priors_match_golden(oracle, golden):
    g = golden["priors_ext"]
    spec = [tuple(s_) for s_ in g["spec"]]
    th = prior_draws(oracle, spec, g["shape"][1])
    assert [float(x).hex() for x in th.ravel()] == g["draws"]
    pri = oracle.make_priors(spec)
    lps = []
    for k in range(len(spec)):
        one = (oracle.Prior * 1)(pri[k])
        for x in list(th[k]) + [0.0, 1.0, 2.5, -1.0, 7.0]:
            lps.append(oracle.lib().kor_prior_logpdf(one, 1, (C.c_double * 1)(float(x))))
    assert [float(x).hex() for x in lps] == g["logpdf"]


def test_socks_cost_against_numpy(oracle):
    """ref test/runtests.jl:34-44 restated with numpy on the spec'd Fisher-Yates draw."""
    L = oracle.lib()
    m = oracle.make_model(oracle.SOCKS, 0, (0, 11), (11,))
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(0, 80)) if trial else 0
        prop = float(rng.uniform(0, 1))
        pid, ep = int(rng.integers(0, 1000)), int(rng.integers(0, 50))
        n_pairs = int(np.rint(prop * math.floor(n / 2)))
        n_odd = n - 2 * n_pairs
        socks = sorted(list(range(1, n_pairs + 1)) * 2 + list(range(n_pairs + 1, n_pairs + n_odd + 1)))
        perm = list(range(n))
        mp = min(n, 11)
        for j in range(mp):
            r = j + L.kor_index(L.kor_stream_word(SEED, 3, pid, ep, j), n - j)
            perm[j], perm[r] = perm[r], perm[j]
        picked = [socks[q] for q in perm[:mp]]
        lu = len(set(picked))
        pairs, odds = mp - lu, lu - (mp - lu)
        th = np.array([float(n), prop])
        c = L.kor_cost(m, SEED, 2, th.ctypes.data_as(C.POINTER(C.c_double)), pid, ep)
        assert c == abs(pairs - 0) + abs(odds - 11)
    th = np.array([-1.0, 0.5])
    assert L.kor_cost(m, SEED, 2, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 0) == math.inf


def test_socks_reference_integration_test(oracle, golden):
    """ref test/runtests.jl:46-74: the posterior of (n_socks, prop_pairs) given (0 pairs, 11 odd): 46.2 and 0.866 from both
    sample(ApproxPosterior(..., 0.1), AIS(500), 5000, ntransitions = 100) and smc(..., nparticles = 5000, alpha = 0.99,
    r_epstol = 0, epstol = 0.01)."""
    M = models(oracle, None)["socks"]
    s = oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](),
                   oracle.smc_config(nparticles=5000, alpha=0.99, r_epstol=0, epstol=0.01), nthreads=8)
    s.run()
    th, X, _, alive = s.state()
    n, p = np.rint(th[0][alive > 0]), th[1][alive > 0]
    assert (X[alive > 0] <= s.scalars()["eps"]).all()
    assert abs(n.mean() - 46.2) / n.std(ddof=1) < 2 and abs(n.mean() - 46.2) < 1.5          # P[1] ≈ 46.2
    assert abs(p.mean() - 0.866) / p.std(ddof=1) < 2 and abs(p.mean() - 0.866) < 0.01       # P[2] ≈ 0.866
    a = oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](),
                   oracle.ais_config(500, 5000, ntransitions=100, scale=0.1, posterior=1), nthreads=8)
    out = a.run_parallel()
    assert (out[0] == np.rint(out[0])).all()                                                 # recorded samples are pushed
    assert abs(out[0].mean() - 46.2) < 1.5 and abs(out[1].mean() - 0.866) < 0.01
    # reduced-size golden runs
    s = oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](),
                   oracle.smc_config(nparticles=500, alpha=0.99, r_epstol=0, epstol=0.01))
    s.run()
    g = golden["smc_socks_500"]
    assert [float(r["eps"]).hex() for r in s.log()] == g["eps"] and [r["n_alive"] for r in s.log()] == g["n_alive"]
    assert float(s.state()[0].sum()).hex() == g["theta_sum"] and s.scalars()["cost_evals"] == g["cost_evals"]


def test_normal_times_discrete_uniform_inference(oracle):
    """ref test/runtests.jl:105-112: Factored(Normal(1,0.5), DiscreteUniform(1,10)), sim((n,du)) = (n*n+du)*(n+randn()*0.01),
    ApproxPosterior(|sim - 5.5|, 0.01), AIS(100), 1000 samples, discard_initial = 10000 -> sim(res) ≈ 5.5."""
    M = models(oracle, None)["noisyprod"]
    a = oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](),
                   oracle.ais_config(100, 1000, discard_initial=10000, scale=0.01, posterior=1), nthreads=8)
    n, du = a.run_sequential()
    assert (du == np.rint(du)).all() and du.min() >= 1 and du.max() <= 10 and len(np.unique(du)) > 3
    sim = (n * n + du) * n
    assert abs(sim.mean() - 5.5) / sim.std(ddof=1) < 2 and abs(sim.mean() - 5.5) < 0.05


# ------------------------------------------------------------------ quantile (Statistics.jl type 7, ref src/smc.jl:134)
def test_quantile_type7_against_numpy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 10, 100, 1001]:
        for p in [0.0, 0.05, 0.5, 0.9, 0.95, 0.99, 1.0]:
            v = rng.normal(size=n)
            q = L.kor_quantile7(v.ctypes.data_as(C.POINTER(C.c_double)), n, p)
            assert abs(q - np.quantile(v, p, method="linear")) <= 1e-14 * max(1, abs(q))
    v = np.array([1.0, 2.0, np.inf, np.inf])
    assert L.kor_quantile7(v.ctypes.data_as(C.POINTER(C.c_double)), 4, 0.5) == np.inf
    v = np.array([3.0, 3.0, 3.0, 3.0, 3.0])
    assert L.kor_quantile7(v.ctypes.data_as(C.POINTER(C.c_double)), 5, 0.95) == 3.0


# ------------------------------------------------------------------ simulators against numpy restatements
def normals_of_stream(oracle, tag, pid, epoch, n):
    L = oracle.lib()
    z0, z1 = C.c_double(), C.c_double()
    out = []
    for b in range((n + 3) // 4):
        w = [L.kor_stream_word(SEED, tag, pid, epoch, 4 * b + q) for q in range(4)]
        L.kor_normal_pair(w[0], w[1], C.byref(z0), C.byref(z1)); out += [z0.value, z1.value]
        L.kor_normal_pair(w[2], w[3], C.byref(z0), C.byref(z1)); out += [z0.value, z1.value]
    return np.array(out[:n])


def test_normal_cost_is_readme_formula(oracle):
    """README.md:46-52: hypot(mean(x)-2.0, (std(x)-0.04)*50) with x = randn(n).*sigma .+ mu, std with n-1."""
    m = oracle.make_model(oracle.NORMAL_MEANSTD, 1000, (2.0, 0.04), (50.0,))
    for pid, (mu, sg) in enumerate([(2.0, 0.04), (1.3, 0.2), (2.9, 0.001)]):
        z = normals_of_stream(oracle, oracle.ST_COST, pid, 4, 1000)
        x = z * sg + mu
        ref = math.hypot(x.mean() - 2.0, (x.std(ddof=1) - 0.04) * 50)
        th = np.array([mu, sg])
        c = oracle.lib().kor_cost(C.byref(m), SEED, 2, th.ctypes.data_as(C.POINTER(C.c_double)), pid, 4)
        assert abs(c - ref) <= 1e-12 * ref


def test_ma2_cost_against_numpy(oracle):
    m = oracle.make_model(oracle.MA2_AUTOCOV, 100, (0.72, 0.2))
    for pid, (t1, t2) in enumerate([(0.6, 0.2), (-1.2, 0.5), (1.9, 0.95)]):
        e = normals_of_stream(oracle, oracle.ST_COST, pid, 1, 102)
        y = e[2:] + t1 * e[1:-1] + t2 * e[:-2]
        tau1, tau2 = (y[1:] * y[:-1]).sum() / 100, (y[2:] * y[:-2]).sum() / 100
        ref = math.hypot(tau1 - 0.72, tau2 - 0.2)
        th = np.array([t1, t2])
        c = oracle.lib().kor_cost(C.byref(m), SEED, 2, th.ctypes.data_as(C.POINTER(C.c_double)), pid, 1)
        assert abs(c - ref) <= 1e-12 * ref
    for t1, t2 in [(2.0, 0.0), (0.0, -1.0), (1.5, -0.6), (-1.5, -0.6)]:  # outside the invertibility triangle
        th = np.array([t1, t2])
        assert oracle.lib().kor_cost(C.byref(m), SEED, 2, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 0) == math.inf


def test_gk_cost_against_numpy(oracle):
    from common import GK_TARGET
    n = 1000
    m = oracle.make_model(oracle.GK_OCTILE, n, GK_TARGET, (0.8,))
    for pid, (A, B, g, k) in enumerate([(3, 1, 2, 0.5), (0.5, 7, 9.5, 0.01), (9, 0.2, 0.0, 3.0)]):
        z = normals_of_stream(oracle, oracle.ST_COST, pid, 2, n)
        x = np.sort(A + B * (1 + 0.8 * np.tanh(g * z / 2)) * (1 + z * z)**k * z)
        q = np.array([x[int(round(i * n / 8 + 1e-9)) - 1] for i in range(1, 8)])
        ref = math.sqrt(((q - np.array(GK_TARGET))**2).sum())
        th = np.array([A, B, g, k], dtype=np.float64)
        c = oracle.lib().kor_cost(C.byref(m), SEED, 4, th.ctypes.data_as(C.POINTER(C.c_double)), pid, 2)
        assert abs(c - ref) <= 1e-10 * ref


def test_lv_properties(oracle):
    M = models(oracle, None)["lv"]
    L = oracle.lib()
    m = M["omodel"]()
    # at the generating parameters the distance is moderate and the event count is in the thousands
    th = np.log([1.0, 0.005, 0.6])
    costs, events = [], []
    for pid in range(40):
        costs.append(L.kor_cost(C.byref(m), SEED, 3, th.ctypes.data_as(C.POINTER(C.c_double)), pid, 0))
        events.append(L.kor_last_events())
    assert 2000 < np.mean(events) < 15000 and np.isfinite(costs).mean() > 0.9
    # zero birth rate: predators and prey die out, few events, finite cost
    th = np.log([1e-9, 1e-9, 5.0])
    c = L.kor_cost(C.byref(m), SEED, 3, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 0)
    assert np.isfinite(c) and L.kor_last_events() <= 200
    # explosive prey growth hits the event cap -> +Inf, exactly max_events events
    th = np.log([50.0, 1e-9, 0.1])
    assert L.kor_cost(C.byref(m), SEED, 3, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 0) == math.inf
    assert L.kor_last_events() == 20000
    # the committed observation vector is one trajectory of this simulator
    out = (C.c_double * 32)()
    L.kor_lv_trajectory.argtypes = [C.POINTER(oracle.Model), C.c_uint64, C.POINTER(C.c_double), C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    th = np.log([1.0, 0.005, 0.6])
    big = M["omodel"](cap=100000)
    assert L.kor_lv_trajectory(C.byref(big), 1, th.ctypes.data_as(C.POINTER(C.c_double)), 0, 0, out) == 0
    from common import LV_TARGET_X, LV_TARGET_Y
    assert [int(v) for v in out] == LV_TARGET_X + LV_TARGET_Y


def test_costs_match_golden(oracle, golden):
    M = models(oracle, None)
    for key, v in golden["costs"].items():
        name, nd = key.rsplit("_", 1)
        th = np.array([fh(x) for x in v["theta"]]).reshape(v["shape"])
        c = oracle.eval_cost(M[name]["omodel"](int(nd)), SEED, th, first_id=5, epoch=9)
        assert [float(x).hex() for x in c] == v["cost"], key


# ------------------------------------------------------------------ smc (ref src/smc.jl:92-206)
def test_smc_argument_checks(oracle):
    pri = oracle.make_priors([("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)])
    m = oracle.make_model(oracle.NORMAL_MEANSTD, 10, (2.0, 0.04), (50.0,))
    cases = [(dict(min_r_ess=0.0), "min_r_ess must be > 0."), (dict(mcmc_retrys=-1), "mcmc_retrys must be >= 0."),
             (dict(alpha=0.0, min_r_ess=0.5), "alpha must be > 0."), (dict(r_epstol=-1.0), "r_epstol must be >= 0"),
             (dict(mcmc_tol=-0.1), "mcmc_tol must be >= 0"), (dict(max_stretch=1.0), "max_stretch must be > 1"),
             (dict(nparticles=6), "nparticles must be >= 7.")]
    for kw, msg in cases:
        with pytest.raises(oracle.OracleError, match=msg.replace(".", r"\.")):
            oracle.Smc(SEED, pri, m, oracle.smc_config(**kw))


def test_smc_defaults_resample_every_iteration_and_tiling(oracle):
    """with alpha=0.95, min_r_ess=alpha^2 the test alpha*ESS <= N*min_r_ess is a knife edge that is true
    after every cut of a fully alive population; the resample is the cyclic tiling idx[k] = idxalive[k mod n]."""
    assert 0.95 * 95 <= 100 * (0.95 * 0.95)
    assert 0.95 * 996147 <= 1048576 * (0.95 * 0.95)
    M = models(oracle, None)["normal"]
    s = oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](50), oracle.smc_config(nparticles=400, max_iterations=6))
    s.init()
    for it in range(6):
        th0, X0, lpi0, _ = s.state()
        s.cut()
        th1, X1, lpi1, alive1 = s.state()
        eps = s.scalars()["eps"]
        idxalive = np.nonzero(X0 < eps)[0]
        assert s.scalars()["n_alive"] == len(idxalive)
        if it == 0:  # no ties yet: trunc(400*.95+.05) = 380 -> exactly 380 costs lie below eps in (v[380], v[381])
            assert len(idxalive) == 380
        assert 0.95 * len(idxalive) <= 400 * (0.95 * 0.95)             # ... and the knife edge fires
        idx = idxalive[np.arange(400) % len(idxalive)]
        assert (X1 == X0[idx]).all() and (th1 == th0[:, idx]).all() and (lpi1 == lpi0[idx]).all() and alive1.all()
        acc, ev, events = s.sweep_range(0, 400)
        s.sweep_commit(acc, ev, events)
        s.finish()
    assert all(r["resampled"] for r in s.log())


def test_smc_matches_golden_run(oracle, golden):
    M = models(oracle, None)["normal"]
    s = oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](100), oracle.smc_config(nparticles=256, max_iterations=12))
    s.run()
    g = golden["smc_normal_256"]
    assert [float(r["eps"]).hex() for r in s.log()] == g["eps"]
    assert [r["n_alive"] for r in s.log()] == g["n_alive"] and [r["accepted"] for r in s.log()] == g["accepted"]
    th, X, _, _ = s.state()
    assert [float(x).hex() for x in th[:, :8].ravel()] == g["theta_first8"] and float(th.sum()).hex() == g["theta_sum"]
    assert s.scalars()["cost_evals"] == g["cost_evals"]


def test_smc_readme_posterior(oracle):
    """README.md:83-84: smc(prior,cost) -> mu = 2.0 +- 0.0062, sigma = 0.0401 +- 0.00081, eps = 0.0111.  Statistical."""
    M = models(oracle, None)["normal"]
    s = oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](1000), oracle.smc_config(nparticles=2000, epstol=0.0111), nthreads=8)
    s.run()
    th, X, _, alive = s.state()
    mu, sg = th[0][alive == 1], th[1][alive == 1]
    assert s.scalars()["eps"] <= 0.0111
    assert abs(mu.mean() - 2.0) < 0.002 and abs(sg.mean() - 0.04) < 0.0005
    assert 0.0005 < sg.std() < 0.002 and 0.0005 < mu.std() < 0.01


def test_smc_deterministic_cost_normal_to_dirac(oracle):
    """ref test/runtests.jl:77-86: prior Normal(1,0.2), cost |mu^2+1-1.5| -> smc(pri,cost,epstol=0.1).P ~ 0.707."""
    pri = oracle.make_priors([("normal", 1, 0.2)])
    m = oracle.make_model(oracle.DETERMINISTIC, 0, (1.5,), (0.0,))
    s = oracle.Smc(SEED, pri, m, oracle.smc_config(epstol=0.1))
    s.run()
    th, _, _, alive = s.state()
    p = th[0][alive == 1]
    assert abs(p.mean() - 0.707) / p.std(ddof=1) < 2  # MonteCarloMeasurements' `≈`


def test_smc_replay_override_is_neutral(oracle):
    """feeding a run's own proposal costs back through the replay hook reproduces the run"""
    M = models(oracle, None)["normal"]
    mk = lambda: oracle.Smc(SEED, oracle.make_priors(M["ospec"]), M["omodel"](50), oracle.smc_config(nparticles=300, alpha=0.8, min_r_ess=0.3, max_iterations=5))
    a, b = mk(), mk()
    a.init(); b.init()
    for _ in range(5):
        a.iterate()
        b.set_cost_override(a.trace()["xp"])
        b.iterate()
        for x, y in zip(a.state(), b.state()):
            assert (np.asarray(x).view(np.uint8) == np.asarray(y).view(np.uint8)).all()
        assert set(np.unique(a.trace()["decision"])) <= {0, 1, 2, 3, 4}


# ------------------------------------------------------------------ AIS (ref src/transition.jl, src/types.jl:51-75)
def test_ais_trace_invariants(oracle):
    M = models(oracle, None)["normal"]
    a = oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](50), oracle.ais_config(40, 1, scale=0.2))
    a.init()
    seen = set()
    for _ in range(30):
        th, lp0, ll0 = a.state()
        a.sweep()
        t = a.trace()
        st = t["move"] == 1
        Z = np.exp(t["corr"][st])                      # corr = (d-1) log Z with d = 2
        assert (Z >= 1 / 3 - 1e-12).all() and (Z < 3 + 1e-12).all()
        assert (t["corr"][~st] == 0).all()
        h = 20
        for i in range(40):                            # partners come from the complementary colour, all distinct
            comp = range(h, 40) if i < h else range(0, h)
            ps = [int(t[k][i]) for k in ("a", "b", "c") if t[k][i] >= 0]
            assert all(p in comp for p in ps) and len(set(ps)) == len(ps)
            assert len(ps) == {1: 1, 2: 2, 3: 3}[int(t["move"][i])]
        inval = t["decision"] == 0                      # invalid proposals consume no accept variate
        assert np.isnan(t["e"][inval]).all() and (t["e"][~inval] > 0).all()
        assert (np.isneginf(t["lp_p"][inval]) | ~np.isfinite(t["ll_p"][inval])).all()
        seen |= set(np.unique(t["move"]))
    assert seen == {1, 2, 3}


def test_ais_matches_golden_run(oracle, golden):
    M = models(oracle, None)["normal"]
    a = oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](100),
                   oracle.ais_config(12, 40, ntransitions=5, discard_initial=3, thinning=2, scale=0.05))
    out = a.run_parallel()
    assert [float(x).hex() for x in out.ravel()] == golden["ais_normal_12"]["samples"]
    assert a.counters() == golden["ais_normal_12"]["counters"]


def test_ais_reference_schedule_counts_and_readme_posterior(oracle):
    """config 1: AIS(10), 1000 samples, ntransitions=100, eps=0.005 (README.md:56-66): 10 + 999*100 = 99 910 transitions,
    each <= 1 cost evaluation; second half of the chain: mu ~ 2.0, sigma ~ 0.04 +- 0.00093."""
    M = models(oracle, None)["normal"]
    a = oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](1000), oracle.ais_config(10, 1000, ntransitions=100, scale=0.005))
    out = a.run_sequential()
    c = a.counters()
    assert 99910 * 0.97 < c["cost_evals"] <= 99910
    mu, sg = out[0][500:], out[1][500:]
    assert abs(mu.mean() - 2.0) < 0.004 and abs(sg.mean() - 0.04) < 0.0005 and 0.0005 < sg.std() < 0.0016


def test_ais_hard_threshold_issue10(oracle):
    """ref test/runtests.jl:177-182: ApproxPosterior(Normal(0,1), x -> abs(x-1.5), 0.01), AIS(20), 100 samples,
    discard_initial = 2000 -> res ≈ 1.5; every retained sample obeys the hard threshold."""
    pri = oracle.make_priors([("normal", 0, 1)])
    m = oracle.make_model(oracle.DETERMINISTIC, 0, (1.5,), (1.0,))
    a = oracle.Ais(SEED, pri, m, oracle.ais_config(20, 100, discard_initial=2000, scale=0.01, posterior=1))
    out = a.run_sequential()[0]
    assert abs(out.mean() - 1.5) / out.std(ddof=1) < 2 and (np.abs(out - 1.5) <= 0.01).all()
    b = oracle.Ais(SEED, pri, m, oracle.ais_config(20, 100, discard_initial=2000, scale=0.01, posterior=1))
    out = b.run_parallel()[0]
    assert abs(out.mean() - 1.5) / out.std(ddof=1) < 2 and (np.abs(out - 1.5) <= 0.01).all()


# ------------------------------------------------------------------ ABCDE / pfilter (ref src/smc.jl:275-428)
def test_abcde_normal_to_dirac_and_invariants(oracle):
    """ref src/smc.jl:352-428 on the cost of test/runtests.jl:77-86 (|mu^2+1-1.5|, prior Normal(1,0.2))."""
    pri = oracle.make_priors([("normal", 1, 0.2)])
    m = oracle.make_model(oracle.DETERMINISTIC, 0, (1.5,), (0.0,))
    r = oracle.abcde(SEED, pri, m, 0.01, nparticles=200, generations=60)
    assert r["reached"] and (r["C"] <= 0.01).all() and r["generations"] == 60
    assert abs(r["theta"].mean() - 0.7071) < 0.005 and 0 < r["nsim"] <= 200 * 60
    e = oracle.abcde(SEED, pri, m, 0.01, nparticles=200, generations=600, earlystop=True)
    assert e["reached"] and e["generations"] < 100 and e["nsim"] < r["nsim"]
    n0 = oracle.abcde(SEED, pri, m, 0.01, nparticles=200, generations=0)          # generations = 0: the prior sample
    assert n0["nsim"] == 0 and not n0["reached"] and abs(n0["theta"].mean() - 1.0) < 0.05
    a = oracle.abcde(SEED, pri, m, 0.01, nparticles=200, generations=60, alpha=0.5)
    assert a["reached"] and (a["theta"] != r["theta"]).any()
    with pytest.raises(oracle.OracleError, match="must be in 0 <="):              # ref :353 @assert
        oracle.abcde(SEED, pri, m, 0.01, alpha=1.0)


def test_abcde_and_pfilter_readme_posterior(oracle):
    """README.md:35-66 model through the two other samplers: posterior around mu = 2, sigma = 0.04."""
    M = models(oracle, None)["normal"]
    r = oracle.abcde(SEED, oracle.make_priors(M["ospec"]), M["omodel"](1000), 0.02, nparticles=300, generations=150, alpha=0.5,
                     nthreads=8)
    assert r["reached"] and abs(r["theta"][0].mean() - 2.0) < 0.01 and abs(r["theta"][1].mean() - 0.04) < 0.004
    f = oracle.pfilter(SEED, oracle.make_priors(M["ospec"]), M["omodel"](1000), 500, nthreads=8)
    assert abs(f["theta"][0].mean() - 2.0) < 0.01 and abs(f["theta"][1].mean() - 0.04) < 0.004
    assert (f["C"] <= f["eps"]).all() and f["iterations"] >= 5 and f["nreps"] >= f["iterations"]


def test_pfilter_semantics(oracle):
    """ref src/smc.jl:275-345: particle-count floor, stop rules, discrete prior through push_p."""
    L = oracle.lib()
    assert L.kor_pfilter_nparticles(5, 2, 0.7) == 13 and L.kor_pfilter_nparticles(100, 2, 0.7) == 100   # ref :276-279
    assert L.kor_pfilter_nparticles(8, 2, 1.0) == 9
    pri = oracle.make_priors([("normal", 1, 0.2)])
    m = oracle.make_model(oracle.DETERMINISTIC, 0, (1.5,), (0.0,))
    r = oracle.pfilter(SEED, pri, m, 500, epstol=0.01)
    assert r["eps"] < 0.01 and (r["C"] <= r["eps"]).all() and abs(r["theta"].mean() - 0.7071) < 0.005   # ref :333
    one = oracle.pfilter(SEED, pri, m, 500, max_iters=1)                                                 # ref :334 iters > max_iters
    assert one["iterations"] == 2
    tiny = oracle.pfilter(SEED, pri, m, 2)
    assert tiny["theta"].shape == (1, 8)
    S = models(oracle, None)["socks"]
    f = oracle.pfilter(SEED, oracle.make_priors(S["ospec"]), S["omodel"](), 1000, max_iters=3, nthreads=8)
    assert (f["theta"][0] == np.rint(f["theta"][0])).all() and (f["C"] <= f["eps"]).all()


def test_abcde_pfilter_match_golden(oracle, golden):
    M = models(oracle, None)
    g = golden["abcde_normal_300"]
    r = oracle.abcde(SEED, oracle.make_priors(M["normal"]["ospec"]), M["normal"]["omodel"](100), 0.05, nparticles=300,
                     generations=25, alpha=0.3)
    assert [float(x).hex() for x in r["theta"].ravel()] == g["theta"] and r["nsim"] == g["nsim"]
    g = golden["pfilter_socks_400"]
    f = oracle.pfilter(SEED, oracle.make_priors(M["socks"]["ospec"]), M["socks"]["omodel"](), 400, max_iters=4)
    assert [float(x).hex() for x in f["theta"].ravel()] == g["theta"] and f["nreps"] == g["nreps"]
    assert float(f["eps"]).hex() == g["eps"]


def test_ais_errors(oracle):
    M = models(oracle, None)["normal"]
    with pytest.raises(oracle.OracleError, match="is insufficient"):
        oracle.Ais(SEED, oracle.make_priors(M["ospec"]), M["omodel"](10), oracle.ais_config(6, 1, scale=1.0))
    # ref src/KissABC.jl:58-59, test/runtests.jl:221-238: every prior draw has infinite cost -> retry budget error
    pri = oracle.make_priors([("uniform", 3, 4), ("uniform", -1, 1)])
    m = oracle.make_model(oracle.MA2_AUTOCOV, 50, (0, 0))
    a = oracle.Ais(SEED, pri, m, oracle.ais_config(10, 1, retry_sampling=5, scale=0.1))
    with pytest.raises(oracle.OracleError, match="Prior leads to"):
        a.init()
