"""Shared fixtures of the parity tests: the same model described once for the oracle and once for the product."""
import numpy as np

SEED = 0x4B49535341424300

import importlib.util as _u
import os as _os

# workload constants come from the product-side config table (pure Python constants, no CUDA involved)
_spec = _u.spec_from_file_location("_kabc_workload_consts", _os.path.join(_os.path.dirname(_os.path.dirname(
    _os.path.abspath(__file__))), "kissabc.jl_b200", "workloads.py"))


def _consts():
    src = open(_spec.origin).read()
    ns = {}
    exec(src[src.index("MA2_TARGET"):src.index("def normal")], ns)  # only the literal tables
    return ns


_c = _consts()
LV_TARGET_X, LV_TARGET_Y, GK_TARGET, MA2_TARGET = _c["LV_TARGET_X"], _c["LV_TARGET_Y"], _c["GK_TARGET"], _c["MA2_TARGET"]


# ref test/runtests.jl:46-50: prior_mu = 30, prior_sd = 15, prior_size = -mu^2/(mu - sd^2); NegativeBinomial(size, size/(mu+size))
SOCKS_R = -30.0 ** 2 / (30.0 - 15.0 ** 2)
SOCKS_P = SOCKS_R / (30.0 + SOCKS_R)


def models(O, k):
    """name -> (oracle prior specs, product prior, oracle model factory(prec), product cost factory(prec))."""
    return {
        "normal": dict(
            d=2,
            ospec=[("uniform", 1, 3), ("truncnormal", 0, 0.1, 0, 100)],
            kprior=lambda: k.Factored(k.Uniform(1, 3), k.Truncated(k.Normal(0, 0.1), 0, 100)),
            omodel=lambda n=1000: O.make_model(O.NORMAL_MEANSTD, n, target=(2.0, 0.04), param=(50.0,)),
            kcost=lambda prec, n=1000: k.NormalMeanStd(n, 2.0, 0.04, 50.0, precision=prec),
        ),
        "ma2": dict(
            d=2,
            ospec=[("uniform", -2, 2), ("uniform", -1, 1)],
            kprior=lambda: k.Factored(k.Uniform(-2, 2), k.Uniform(-1, 1)),
            omodel=lambda n=100: O.make_model(O.MA2_AUTOCOV, n, target=MA2_TARGET),
            kcost=lambda prec, n=100: k.MA2(n, MA2_TARGET, precision=prec),
        ),
        "gk": dict(
            d=4,
            ospec=[("uniform", 0, 10)] * 4,
            kprior=lambda: k.Factored(*[k.Uniform(0, 10)] * 4),
            omodel=lambda n=10000: O.make_model(O.GK_OCTILE, n, target=GK_TARGET, param=(0.8,)),
            kcost=lambda prec, n=10000: k.GandK(n, GK_TARGET, 0.8, precision=prec),
        ),
        "lv": dict(
            d=3,
            ospec=[("uniform", -2, 1), ("uniform", -7, -4), ("uniform", -2, 1)],
            kprior=lambda: k.Factored(k.Uniform(-2, 1), k.Uniform(-7, -4), k.Uniform(-2, 1)),
            omodel=lambda n=0, cap=20000: O.make_model(O.LV_SSA, 0, target=LV_TARGET_X + LV_TARGET_Y,
                                                       param=(50, 100, 30, 16, cap)),
            kcost=lambda prec, n=0, cap=20000: k.LotkaVolterra(LV_TARGET_X + LV_TARGET_Y, 50, 100, 30, cap, precision=prec),
        ),
        # ref test/runtests.jl:34-56 (socks of Karl Broman: NegativeBinomial x Beta prior, discrete first component)
        "socks": dict(
            d=2,
            ospec=[("negbin", SOCKS_R, SOCKS_P), ("beta", 15, 2)],
            kprior=lambda: k.Factored(k.NegativeBinomial(SOCKS_R, SOCKS_P), k.Beta(15, 2)),
            omodel=lambda n=0: O.make_model(O.SOCKS, 0, target=(0, 11), param=(11,)),
            kcost=lambda prec, n=0: k.Socks((0, 11), 11),
        ),
        # ref test/runtests.jl:105-112 (Normal x DiscreteUniform prior, noisy product simulator)
        "noisyprod": dict(
            d=2,
            ospec=[("normal", 1, 0.5), ("duniform", 1, 10)],
            kprior=lambda: k.Factored(k.Normal(1, 0.5), k.DiscreteUniform(1, 10)),
            omodel=lambda n=0: O.make_model(O.DETERMINISTIC, 0, target=(5.5,), param=(2.0, 0.01)),
            kcost=lambda prec, n=0: k.NoisyProduct(5.5, 0.01),
        ),
    }


def prior_draws(O, ospec, n, seed=SEED, epoch=7):
    """n prior draws (d x n, SoA) from the oracle's prior sampler."""
    import ctypes as C
    pri = O.make_priors(ospec)
    d = len(ospec)
    th = np.empty((d, n))
    buf = (C.c_double * d)()
    for i in range(n):
        O.lib().kor_prior_sample(seed, pri, d, i, epoch, buf)
        th[:, i] = list(buf)
    return th
