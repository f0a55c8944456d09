"""ctypes binding of the CPU ORACLE (oracle/kabc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

The oracle restates the reference algorithm (KissABC.jl 3.0.1: src/transition.jl,
src/types.jl:51-75, src/priors.jl:30-43, src/KissABC.jl:35-80, src/smc.jl:92-206)
over the canonical Philox variate source (DESIGN.md "Variate spec").
Stream-level parity with Julia is UNPINNED (no Julia here, no golden streams upstream).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class Prior(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("p0", C.c_double), ("p1", C.c_double),
                ("lo", C.c_double), ("hi", C.c_double)]


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("precision", C.c_int32), ("n_draws", C.c_int32), ("n_target", C.c_int32),
                ("target", C.c_double * 32), ("param", C.c_double * 8)]


class SmcConfig(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("alpha", C.c_double), ("mcmc_retrys", C.c_int64),
                ("mcmc_tol", C.c_double), ("epstol", C.c_double), ("r_epstol", C.c_double),
                ("min_r_ess", C.c_double), ("max_stretch", C.c_double), ("verbose", C.c_int32),
                ("max_iterations", C.c_int32)]


class AisConfig(C.Structure):
    _fields_ = [("nwalkers", C.c_int64), ("nsamples", C.c_int64), ("ntransitions", C.c_int64),
                ("discard_initial", C.c_int64), ("thinning", C.c_int64), ("retry_sampling", C.c_int64),
                ("scale", C.c_double), ("posterior", C.c_int32), ("_pad", C.c_int32)]


class AbcdeConfig(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("generations", C.c_int64), ("eps_target", C.c_double), ("alpha", C.c_double),
                ("proposal_width", C.c_double), ("earlystop", C.c_int32), ("_pad", C.c_int32)]


class PfilterConfig(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("q", C.c_double), ("eff_tol", C.c_double), ("epstol", C.c_double),
                ("proposal_width", C.c_double), ("max_iters", C.c_int64)]


class SmcLog(C.Structure):
    _fields_ = [("iteration", C.c_int64), ("eps", C.c_double), ("n_alive", C.c_int64), ("flag", C.c_int32),
                ("resampled", C.c_int32), ("accepted", C.c_int64), ("cost_evals", C.c_int64),
                ("sweeps", C.c_int64)]


UNIFORM, NORMAL, TRUNC_NORMAL, BETA, NEG_BINOMIAL, DISCRETE_UNIFORM = 0, 1, 2, 3, 4, 5
NORMAL_MEANSTD, MA2_AUTOCOV, GK_OCTILE, LV_SSA, DETERMINISTIC, SOCKS = 0, 1, 2, 3, 4, 5
ST_PRIOR, ST_PROPOSE, ST_COST, ST_ACCEPT, ST_COST_INIT = 1, 2, 3, 4, 5


def _has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return False


def build(force: bool = False) -> None:
    """Compile the oracle with gcc (building the checker is not using it)."""
    want = [os.path.join(_HERE, n) for n in ("libkabc_oracle.so", "libkabc_oracle_fma.so")]
    src = [os.path.join(_HERE, n) for n in ("kabc_oracle.c", "kabc_oracle.h")]
    stale = force or any(not os.path.exists(w) or os.path.getmtime(w) < max(os.path.getmtime(s) for s in src)
                         for w in want)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "CC=gcc"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    name = "libkabc_oracle_fma.so" if _has_fma() else "libkabc_oracle.so"
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    dp, u8p, i64p = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.kor_last_error.restype = C.c_char_p
    L.kor_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    for f in (L.kor_log, L.kor_exp, L.kor_lgamma):
        f.argtypes = [C.c_double]
        f.restype = C.c_double
    L.kor_sincos2pi.argtypes = [C.c_double, dp, dp]
    L.kor_u01.argtypes = [C.c_uint32]
    L.kor_u01.restype = C.c_double
    L.kor_index.argtypes = [C.c_uint32, C.c_uint32]
    L.kor_index.restype = C.c_uint32
    L.kor_normal_pair.argtypes = [C.c_uint32, C.c_uint32, dp, dp]
    L.kor_stream_word.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.kor_stream_word.restype = C.c_uint32
    L.kor_prior_logpdf.argtypes = [C.POINTER(Prior), C.c_int, dp]
    L.kor_prior_logpdf.restype = C.c_double
    L.kor_prior_sample.argtypes = [C.c_uint64, C.POINTER(Prior), C.c_int, C.c_uint32, C.c_uint32, dp]
    L.kor_push_p.argtypes = [C.POINTER(Prior), C.c_int, dp, dp]
    L.kor_push_p.restype = None
    L.kor_cost.argtypes = [C.POINTER(Model), C.c_uint64, C.c_int, dp, C.c_uint32, C.c_uint32]
    L.kor_cost.restype = C.c_double
    L.kor_last_events.restype = C.c_int64
    L.kor_eval_cost.argtypes = [C.POINTER(Model), C.c_uint64, C.c_int, dp, C.c_int64, C.c_uint32, C.c_uint32, dp, C.c_int]
    L.kor_quantile7.argtypes = [dp, C.c_int64, C.c_double]
    L.kor_quantile7.restype = C.c_double
    L.kor_smc_create.argtypes = [C.c_uint64, C.POINTER(Prior), C.c_int, C.POINTER(Model), C.POINTER(SmcConfig), C.c_int, C.POINTER(vp)]
    L.kor_smc_destroy.argtypes = [vp]
    L.kor_smc_init.argtypes = [vp]
    L.kor_smc_iterate.argtypes = [vp, C.POINTER(C.c_int)]
    L.kor_smc_run.argtypes = [vp]
    L.kor_smc_cut.argtypes = [vp]
    L.kor_smc_sweep_range.argtypes = [vp, C.c_int64, C.c_int64, i64p, i64p, i64p]
    L.kor_smc_sweep_commit.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
    L.kor_smc_finish.argtypes = [vp, C.POINTER(C.c_int)]
    L.kor_smc_set_cost_override.argtypes = [vp, dp]
    L.kor_smc_set_serial.argtypes = [vp]
    L.kor_ais_set_serial.argtypes = [vp]
    L.kor_smc_get_state.argtypes = [vp, dp, dp, dp, u8p]
    L.kor_smc_set_state.argtypes = [vp, dp, dp, dp, u8p]
    L.kor_smc_get_scalars.argtypes = [vp, dp, C.POINTER(C.c_int32), i64p, i64p, i64p, i64p, i64p]
    L.kor_smc_get_log.argtypes = [vp, C.POINTER(SmcLog), C.c_int64]
    L.kor_smc_get_log.restype = C.c_int64
    L.kor_smc_get_trace.argtypes = [vp, i64p, i64p, dp, dp, dp, dp, u8p, dp]
    L.kor_abcde_run.argtypes = [C.c_uint64, C.POINTER(Prior), C.c_int, C.POINTER(Model), C.POINTER(AbcdeConfig), C.c_int, dp, dp,
                                C.POINTER(C.c_int32), i64p, i64p]
    L.kor_pfilter_nparticles.argtypes = [C.c_int64, C.c_int, C.c_double]
    L.kor_pfilter_nparticles.restype = C.c_int64
    L.kor_pfilter_run.argtypes = [C.c_uint64, C.POINTER(Prior), C.c_int, C.POINTER(Model), C.POINTER(PfilterConfig), C.c_int, dp, dp,
                                  dp, i64p, i64p, i64p]
    L.kor_ais_create.argtypes = [C.c_uint64, C.POINTER(Prior), C.c_int, C.POINTER(Model), C.POINTER(AisConfig), C.c_int, C.POINTER(vp)]
    L.kor_ais_destroy.argtypes = [vp]
    L.kor_ais_init.argtypes = [vp]
    L.kor_ais_transition.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_uint32]
    L.kor_ais_sweep.argtypes = [vp]
    L.kor_ais_run_sequential.argtypes = [vp, dp]
    L.kor_ais_run_parallel.argtypes = [vp, dp]
    L.kor_ais_get_state.argtypes = [vp, dp, dp, dp]
    L.kor_ais_set_state.argtypes = [vp, dp, dp, dp]
    L.kor_ais_get_counters.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.kor_ais_get_trace.argtypes = [vp, u8p, i64p, i64p, i64p, dp, dp, dp, dp, dp, u8p]
    _lib = L
    return L


class OracleError(RuntimeError):
    pass


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def make_priors(specs):
    """specs: list of ("uniform",a,b) | ("normal",mu,sigma) | ("truncnormal",mu,sigma,lo,hi) | ("beta",a,b) |
    ("negbin",r,p) | ("duniform",a,b)."""
    arr = (Prior * len(specs))()
    for k, s in enumerate(specs):
        kind = s[0]
        if kind == "uniform":
            arr[k] = Prior(UNIFORM, 0, float(s[1]), float(s[2]), float(s[1]), float(s[2]))
        elif kind == "normal":
            arr[k] = Prior(NORMAL, 0, float(s[1]), float(s[2]), -np.inf, np.inf)
        elif kind == "truncnormal":
            arr[k] = Prior(TRUNC_NORMAL, 0, float(s[1]), float(s[2]), float(s[3]), float(s[4]))
        elif kind == "beta":
            arr[k] = Prior(BETA, 0, float(s[1]), float(s[2]), 0.0, 1.0)
        elif kind == "negbin":
            arr[k] = Prior(NEG_BINOMIAL, 0, float(s[1]), float(s[2]), 0.0, np.inf)
        elif kind == "duniform":
            arr[k] = Prior(DISCRETE_UNIFORM, 0, float(s[1]), float(s[2]), float(s[1]), float(s[2]))
        else:
            raise ValueError(kind)
    return arr


def make_model(kind, n_draws=0, target=(), param=()):
    m = Model()
    m.kind, m.precision, m.n_draws, m.n_target = kind, 0, int(n_draws), len(target)
    for i, t in enumerate(target):
        m.target[i] = float(t)
    for i, p in enumerate(param):
        m.param[i] = float(p)
    return m


def smc_config(nparticles=100, alpha=0.95, mcmc_retrys=0, mcmc_tol=0.015, epstol=0.0, r_epstol=None,
               min_r_ess=None, max_stretch=2.0, verbose=False, max_iterations=0):
    """Defaults of ref src/smc.jl:95-105."""
    if r_epstol is None:
        r_epstol = (1 - alpha) ** 1.5 / 50
    if min_r_ess is None:
        min_r_ess = alpha * alpha  # alpha^2 in Julia == alpha*alpha (literal_pow)
    return SmcConfig(int(nparticles), float(alpha), int(mcmc_retrys), float(mcmc_tol), float(epstol),
                     float(r_epstol), float(min_r_ess), float(max_stretch), int(verbose), int(max_iterations))


def ais_config(nwalkers, nsamples, ntransitions=1, discard_initial=0, thinning=1, retry_sampling=100, scale=1.0,
               posterior=0):
    return AisConfig(int(nwalkers), int(nsamples), int(ntransitions), int(discard_initial), int(thinning),
                     int(retry_sampling), float(scale), int(posterior), 0)


def abcde(seed, priors, model, eps_target, nparticles=50, generations=20, alpha=0.0, earlystop=False, proposal_width=1.0,
          nthreads=1):
    """ref src/smc.jl:352 ABCDE(prior, cost, eps_target; nparticles=50, generations=20, alpha=0, earlystop=false, ...)."""
    d, N = len(priors), int(nparticles)
    cfg = AbcdeConfig(N, int(generations), float(eps_target), float(alpha), float(proposal_width), int(bool(earlystop)), 0)
    th, cost = np.empty((d, N)), np.empty(N)
    reached, nsim, gens = C.c_int32(0), C.c_int64(0), C.c_int64(0)
    if lib().kor_abcde_run(seed, priors, d, C.byref(model), C.byref(cfg), nthreads, _dp(th), _dp(cost), C.byref(reached),
                           C.byref(nsim), C.byref(gens)):
        raise OracleError(lib().kor_last_error().decode())
    return dict(theta=th, C=cost, reached=bool(reached.value), nsim=nsim.value, generations=gens.value)


def pfilter(seed, priors, model, nparticles, q=0.7, eff_tol=0.1, epstol=-np.inf, max_iters=0, proposal_width=0.75, nthreads=1):
    """ref src/smc.jl:275 pfilter(prior, cost, N; q=0.7, eff_tol=0.1, epstol=-Inf, max_iters=Inf, proposal_width=0.75)."""
    d = len(priors)
    N = int(lib().kor_pfilter_nparticles(int(nparticles), d, float(q)))
    cfg = PfilterConfig(int(nparticles), float(q), float(eff_tol), float(epstol), float(proposal_width), int(max_iters))
    th, cost = np.empty((d, N)), np.empty(N)
    eps, iters, nreps, evals = C.c_double(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    if lib().kor_pfilter_run(seed, priors, d, C.byref(model), C.byref(cfg), nthreads, _dp(th), _dp(cost), C.byref(eps),
                             C.byref(iters), C.byref(nreps), C.byref(evals)):
        raise OracleError(lib().kor_last_error().decode())
    return dict(theta=th, C=cost, eps=eps.value, iterations=iters.value, nreps=nreps.value, cost_evals=evals.value)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().kor_philox4x32_10(c, k, o)
    return tuple(int(x) for x in o)


def eval_cost(model, seed, theta_soa, first_id=0, epoch=0, nthreads=1):
    theta_soa = np.ascontiguousarray(theta_soa, dtype=np.float64)
    d, n = theta_soa.shape
    out = np.empty(n, dtype=np.float64)
    lib().kor_eval_cost(C.byref(model), seed, d, _dp(theta_soa), n, first_id, epoch, _dp(out), nthreads)
    return out


class Smc:
    def __init__(self, seed, priors, model, cfg, nthreads=1):
        self.L = lib()
        self.d = len(priors)
        self.N = int(cfg.nparticles)
        self.h = C.c_void_p()
        self._keep = None
        if self.L.kor_smc_create(seed, priors, self.d, C.byref(model), C.byref(cfg), nthreads, C.byref(self.h)):
            raise OracleError(self.L.kor_last_error().decode())

    def close(self):
        if self.h:
            self.L.kor_smc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_serial(self):
        """one word stream in the reference's consumption order (julia/PhiloxRNG.jl replays); call before init()"""
        self.L.kor_smc_set_serial(self.h)

    def init(self):
        if self.L.kor_smc_init(self.h):
            raise OracleError(self.L.kor_last_error().decode())

    def iterate(self):
        stop = C.c_int(0)
        if self.L.kor_smc_iterate(self.h, C.byref(stop)):
            raise OracleError(self.L.kor_last_error().decode())
        return stop.value

    def run(self):
        if self.L.kor_smc_run(self.h):
            raise OracleError(self.L.kor_last_error().decode())

    # kor_smc_iterate in parts (sharded-schedule emulation)
    def cut(self):
        if self.L.kor_smc_cut(self.h):
            raise OracleError(self.L.kor_last_error().decode())

    def sweep_range(self, lo, hi):
        v = [C.c_int64() for _ in range(3)]
        self.L.kor_smc_sweep_range(self.h, lo, hi, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def sweep_commit(self, acc, evals, events):
        return bool(self.L.kor_smc_sweep_commit(self.h, acc, evals, events))

    def finish(self):
        stop = C.c_int(0)
        self.L.kor_smc_finish(self.h, C.byref(stop))
        return stop.value

    def set_cost_override(self, xp):
        if xp is None:
            self._keep = None
            self.L.kor_smc_set_cost_override(self.h, None)
        else:
            self._keep = np.ascontiguousarray(xp, dtype=np.float64)
            self.L.kor_smc_set_cost_override(self.h, _dp(self._keep))

    def state(self):
        th = np.empty((self.d, self.N)); X = np.empty(self.N); lpi = np.empty(self.N)
        alive = np.empty(self.N, dtype=np.uint8)
        self.L.kor_smc_get_state(self.h, _dp(th), _dp(X), _dp(lpi), _bp(alive))
        return th, X, lpi, alive

    def set_state(self, th, X, lpi, alive):
        th = np.ascontiguousarray(th, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
        lpi = np.ascontiguousarray(lpi, dtype=np.float64); alive = np.ascontiguousarray(alive, dtype=np.uint8)
        self.L.kor_smc_set_state(self.h, _dp(th), _dp(X), _dp(lpi), _bp(alive))

    def scalars(self):
        eps = C.c_double(); flag = C.c_int32(); v = [C.c_int64() for _ in range(5)]
        self.L.kor_smc_get_scalars(self.h, C.byref(eps), C.byref(flag), *[C.byref(x) for x in v])
        return dict(eps=eps.value, flag=flag.value, iteration=v[0].value, n_alive=v[1].value,
                    accepted=v[2].value, cost_evals=v[3].value, next_epoch=v[4].value)

    def log(self):
        n = self.L.kor_smc_get_log(self.h, None, 0)
        buf = (SmcLog * max(n, 1))()
        self.L.kor_smc_get_log(self.h, buf, n)
        return [dict((f, getattr(buf[i], f)) for f, _ in SmcLog._fields_) for i in range(n)]

    def trace(self):
        N, d = self.N, self.d
        a = np.empty(N, np.int64); b = np.empty(N, np.int64)
        z = np.empty(N); lprob = np.empty(N); lpip = np.empty(N); xp = np.empty(N)
        dec = np.empty(N, np.uint8); thp = np.empty((d, N))
        self.L.kor_smc_get_trace(self.h, _ip(a), _ip(b), _dp(z), _dp(lprob), _dp(lpip), _dp(xp), _bp(dec), _dp(thp))
        return dict(a=a, b=b, z=z, lprob=lprob, lpi_p=lpip, xp=xp, decision=dec, theta_p=thp)


class Ais:
    def __init__(self, seed, priors, model, cfg, nthreads=1):
        self.L = lib()
        self.d = len(priors)
        self.N = int(cfg.nwalkers)
        self.Ns = int(cfg.nsamples)
        self.h = C.c_void_p()
        if self.L.kor_ais_create(seed, priors, self.d, C.byref(model), C.byref(cfg), nthreads, C.byref(self.h)):
            raise OracleError(self.L.kor_last_error().decode())

    def close(self):
        if self.h:
            self.L.kor_ais_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise OracleError(self.L.kor_last_error().decode())

    def set_serial(self):
        self.L.kor_ais_set_serial(self.h)

    def init(self):
        self._chk(self.L.kor_ais_init(self.h))

    def transition(self, i, lo, n, epoch):
        return self.L.kor_ais_transition(self.h, i, lo, n, epoch)

    def sweep(self):
        self._chk(self.L.kor_ais_sweep(self.h))

    def run_sequential(self):
        out = np.empty((self.d, self.Ns))
        self._chk(self.L.kor_ais_run_sequential(self.h, _dp(out)))
        return out

    def run_parallel(self):
        out = np.empty((self.d, self.Ns))
        self._chk(self.L.kor_ais_run_parallel(self.h, _dp(out)))
        return out

    def state(self):
        th = np.empty((self.d, self.N)); lp = np.empty(self.N); ll = np.empty(self.N)
        self.L.kor_ais_get_state(self.h, _dp(th), _dp(lp), _dp(ll))
        return th, lp, ll

    def set_state(self, th, lp, ll):
        th = np.ascontiguousarray(th, dtype=np.float64)
        lp = np.ascontiguousarray(lp, dtype=np.float64); ll = np.ascontiguousarray(ll, dtype=np.float64)
        self.L.kor_ais_set_state(self.h, _dp(th), _dp(lp), _dp(ll))

    def counters(self):
        v = [C.c_int64() for _ in range(4)]
        self.L.kor_ais_get_counters(self.h, *[C.byref(x) for x in v])
        return dict(cost_evals=v[0].value, accepted=v[1].value, sweeps=v[2].value, retries=v[3].value)

    def trace(self):
        N, d = self.N, self.d
        move = np.empty(N, np.uint8); dec = np.empty(N, np.uint8)
        a = np.empty(N, np.int64); b = np.empty(N, np.int64); c = np.empty(N, np.int64)
        corr = np.empty(N); thp = np.empty((d, N)); lpp = np.empty(N); llp = np.empty(N); e = np.empty(N)
        self.L.kor_ais_get_trace(self.h, _bp(move), _ip(a), _ip(b), _ip(c), _dp(corr), _dp(thp), _dp(lpp), _dp(llp), _dp(e), _bp(dec))
        return dict(move=move, a=a, b=b, c=c, corr=corr, theta_p=thp, lp_p=lpp, ll_p=llp, e=e, decision=dec)
