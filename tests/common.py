"""Shared fixtures of the parity tests: the same model described once for the oracle and once for the product."""
import numpy as np

SEED = 0x4B49535341424300

# LV observations: a fixed synthetic trajectory on the 16-point grid (generated once by the oracle simulator
# at theta = log(1, 0.005, 0.6), oracle seed 1, id 0, epoch 0 (kor_lv_trajectory); kept literal so oracle and device see the same targets)
LV_TARGET_X = [107, 228, 113, 38, 50, 115, 348, 66, 15, 32, 97, 313, 129, 29, 38, 132]
LV_TARGET_Y = [87, 149, 330, 209, 100, 65, 153, 484, 231, 99, 71, 116, 417, 268, 111, 62]
GK_TARGET = [2.3943, 2.5691, 2.7479, 2.9994, 3.4156, 4.1956, 5.8946]  # octiles of g-and-k(3,1,2,0.5), c=0.8
MA2_TARGET = [0.72, 0.2]  # E[tau1], E[tau2] at theta = (0.6, 0.2): th1 + th1 th2, th2


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
