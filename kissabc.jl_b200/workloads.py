"""The benchmark configurations of BASELINE.json as product-side descriptors."""
from .api import Deterministic, Factored, GandK, LotkaVolterra, MA2, NormalMeanStd, Normal, Truncated, Uniform

# observed summaries (fixed synthetic data; no external datasets)
MA2_TARGET = [0.72, 0.2]  # E[tau1], E[tau2] at theta = (0.6, 0.2): th1 + th1 th2, th2
GK_TARGET = [2.3943, 2.5691, 2.7479, 2.9994, 3.4156, 4.1956, 5.8946]  # octiles of g-and-k(3,1,2,0.5), c = 0.8
# one Gillespie trajectory at c = (1, 0.005, 0.6), X0 = 50, Y0 = 100, observed at t = 1.875 g, g = 1..16
LV_TARGET_X = [107, 228, 113, 38, 50, 115, 348, 66, 15, 32, 97, 313, 129, 29, 38, 132]
LV_TARGET_Y = [87, 149, 330, 209, 100, 65, 153, 484, 231, 99, 71, 116, 417, 268, 111, 62]
LV_MAX_EVENTS = 20000


def normal(prec="f32", n=1000):
    """configs[0]/[1]: README.md:35-52 normal model."""
    return Factored(Uniform(1, 3), Truncated(Normal(0, 0.1), 0, 100)), NormalMeanStd(n, 2.0, 0.04, 50.0, precision=prec)


def ma2(prec="f32", n=100):
    """configs[2]: MA(2), n = 100, autocovariance distance."""
    return Factored(Uniform(-2, 2), Uniform(-1, 1)), MA2(n, MA2_TARGET, precision=prec)


def gk(prec="f32", n=10000):
    """configs[3]: g-and-k, 4 parameters, 10^4 draws, octile distance."""
    return Factored(*[Uniform(0, 10)] * 4), GandK(n, GK_TARGET, 0.8, precision=prec)


def lv(prec="f32", n=0, cap=LV_MAX_EVENTS):
    """configs[4]: stochastic Lotka-Volterra, uniform priors on the log rates."""
    del n
    return (Factored(Uniform(-2, 1), Uniform(-7, -4), Uniform(-2, 1)),
            LotkaVolterra(LV_TARGET_X + LV_TARGET_Y, 50, 100, 30, cap, precision=prec))


def null(prec="f64", n=0):
    """SURVEY.md 8(d) sanity sweep: smc with a (nearly) null simulator, |theta - 1.5| -- what is left is the state traffic of an
    iteration (quantile, cut, table, gathers, accept), to be read against the measured HBM bandwidth."""
    del prec, n
    return Factored(Uniform(0, 3)), Deterministic(1, 1.5)


WORKLOADS = {"normal_smc": normal, "ma2_smc": ma2, "gk_ais": gk, "lv_smc": lv, "null_smc": null}
