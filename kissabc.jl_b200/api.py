"""Host-side mirror of the KissABC.jl interface for the device hot path.

Same names, argument meaning and error behaviour as the reference (KissABC.jl 3.0.1):

    prior = Factored(Uniform(1, 3), Truncated(Normal(0, 0.1), 0, 100))      # src/priors.jl:10-13
    cost  = NormalMeanStd(1000, mean=2.0, std=0.04)                        # replaces the `cost` closure
    res   = smc(prior, cost, nparticles=2**20)                              # src/smc.jl:92-206 -> (P, C, eps)
    post  = ApproxKernelizedPosterior(prior, cost, 0.005)                   # src/types.jl:40-49
    res   = sample(post, AIS(10), 1000, ntransitions=100)                   # src/KissABC.jl:35-94

The `cost` closure of the reference is an arbitrary Julia callable; on the device it is a registered
simulator+distance descriptor (`DeviceCost`).  Everything below is argument marshalling around the C ABI of
libkissabc_cuda.so (include/kissabc_cuda.h): there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import numpy as np

from . import _capi as K
from ._capi import KissABCError

# --------------------------------------------------------------------------------------- distributions


class UnivariateDistribution:
    def _pod(self) -> K.PriorT:
        raise NotImplementedError


@dataclass(frozen=True)
class Uniform(UnivariateDistribution):
    a: float = 0.0
    b: float = 1.0

    def _pod(self):
        return K.PriorT(K.PRIOR_UNIFORM, 0, float(self.a), float(self.b), float(self.a), float(self.b))


@dataclass(frozen=True)
class Normal(UnivariateDistribution):
    mu: float = 0.0
    sigma: float = 1.0

    def _pod(self):
        return K.PriorT(K.PRIOR_NORMAL, 0, float(self.mu), float(self.sigma), -math.inf, math.inf)


@dataclass(frozen=True)
class Truncated(UnivariateDistribution):
    """Truncated(Normal(mu, sigma), lo, hi) -- the only truncated law the device path registers."""
    base: Normal
    lo: float
    hi: float

    def __post_init__(self):
        if not isinstance(self.base, Normal):
            raise KissABCError(K.ERR_INVALID_ARG, "only Truncated(Normal(...), lo, hi) is registered on the device")

    def _pod(self):
        return K.PriorT(K.PRIOR_TRUNC_NORMAL, 0, float(self.base.mu), float(self.base.sigma), float(self.lo), float(self.hi))


@dataclass(frozen=True)
class Beta(UnivariateDistribution):
    """Beta(alpha, beta), ref test/runtests.jl:51, examples/example_n2.jl:28."""
    alpha: float = 1.0
    beta: float = 1.0

    def _pod(self):
        return K.PriorT(K.PRIOR_BETA, 0, float(self.alpha), float(self.beta), 0.0, 1.0)


@dataclass(frozen=True)
class NegativeBinomial(UnivariateDistribution):
    """NegativeBinomial(r, p) (Distributions.jl parametrisation: failures before the r-th success), discrete:
    particles are rounded (push_p, ref src/types.jl:32) before the prior density and the cost see them."""
    r: float = 1.0
    p: float = 0.5

    def _pod(self):
        return K.PriorT(K.PRIOR_NEG_BINOMIAL, 0, float(self.r), float(self.p), 0.0, math.inf)


@dataclass(frozen=True)
class DiscreteUniform(UnivariateDistribution):
    """DiscreteUniform(a, b) on the integers a..b, discrete (push_p, ref src/types.jl:32)."""
    a: int = 0
    b: int = 1

    def _pod(self):
        return K.PriorT(K.PRIOR_DISCRETE_UNIFORM, 0, float(self.a), float(self.b), float(self.a), float(self.b))


class Factored:
    """Product prior, ref src/priors.jl:10-13.  `length(Factored)` = number of components (:49)."""

    def __init__(self, *args: UnivariateDistribution):
        if not args:
            raise KissABCError(K.ERR_INVALID_ARG, "Factored needs at least one distribution")
        for a in args:
            if not isinstance(a, UnivariateDistribution):
                raise KissABCError(K.ERR_INVALID_ARG, f"Factored components must be univariate distributions, got {a!r}")
        self.p = tuple(args)

    def __len__(self):
        return len(self.p)

    def _pods(self):
        arr = (K.PriorT * len(self.p))()
        for k, q in enumerate(self.p):
            arr[k] = q._pod()
        return arr


def _as_factored(prior) -> Factored:
    if isinstance(prior, Factored):
        return prior
    if isinstance(prior, UnivariateDistribution):
        return Factored(prior)
    raise KissABCError(K.ERR_INVALID_ARG, f"unsupported prior {prior!r}")


# --------------------------------------------------------------------------------------- device costs

_PREC = {"f64": K.F64, "f32": K.F32_ACC64, K.F64: K.F64, K.F32_ACC64: K.F32_ACC64}


class DeviceCost:
    """A registered device simulator + distance: what replaces `cost(x)` (src/types.jl:55, src/smc.jl:123,176)."""
    kind: int = -1
    ndim: int = 0

    def __init__(self, n_draws: int, target: Sequence[float], param: Sequence[float], precision="f32"):
        if precision not in _PREC:
            raise KissABCError(K.ERR_INVALID_ARG, f"precision must be 'f64' or 'f32', got {precision!r}")
        self.n_draws = int(n_draws)
        self.target = [float(t) for t in target]
        self.param = [float(p) for p in param]
        self.precision = _PREC[precision]

    def _pod(self) -> K.ModelT:
        m = K.ModelT()
        m.kind, m.precision, m.n_draws, m.n_target = self.kind, self.precision, self.n_draws, len(self.target)
        for i, t in enumerate(self.target):
            m.target[i] = t
        for i, p in enumerate(self.param):
            m.param[i] = p
        return m


class NormalMeanStd(DeviceCost):
    """README.md:35-52: x = randn(n).*sigma .+ mu; hypot(mean(x)-mean_y, (std(x)-std_y)*weight)."""
    kind, ndim = K.MODEL_NORMAL_MEANSTD, 2

    def __init__(self, n_draws=1000, mean=2.0, std=0.04, weight=50.0, precision="f32"):
        super().__init__(n_draws, (mean, std), (weight,), precision)


class MA2(DeviceCost):
    """MA(2) series of n points; Euclidean distance between lag-1/lag-2 autocovariances and `target`."""
    kind, ndim = K.MODEL_MA2_AUTOCOV, 2

    def __init__(self, n=100, target=(0.0, 0.0), precision="f32"):
        super().__init__(n, tuple(target), (), precision)


class GandK(DeviceCost):
    """g-and-k distribution, n draws, Euclidean distance between the 7 octiles and `target` (7 values)."""
    kind, ndim = K.MODEL_GK_OCTILE, 4

    def __init__(self, n=10000, target=(0.0,) * 7, c=0.8, precision="f32"):
        if len(target) != 7:
            raise KissABCError(K.ERR_INVALID_ARG, "g-and-k needs 7 target octiles")
        super().__init__(n, tuple(target), (c,), precision)


class LotkaVolterra(DeviceCost):
    """Stochastic Lotka-Volterra (Gillespie); theta = log rates; RMS distance to observations on a grid."""
    kind, ndim = K.MODEL_LV_SSA, 3

    def __init__(self, target, x0=50.0, y0=100.0, T=30.0, max_events=50000, precision="f32"):
        target = list(target)
        if len(target) % 2 or not (2 <= len(target) <= 32):
            raise KissABCError(K.ERR_INVALID_ARG, "Lotka-Volterra target = [X(t_1..t_G), Y(t_1..t_G)], G <= 16")
        super().__init__(0, target, (x0, y0, T, len(target) // 2, max_events), precision)


class Deterministic(DeviceCost):
    """The deterministic costs of the reference's own tests: variant 0 = |theta^2+1-t| (test/runtests.jl:77-86),
    variant 1 = |theta - t| (test/runtests.jl:177-182)."""
    kind, ndim = K.MODEL_DETERMINISTIC, 1

    def __init__(self, variant=0, target=1.5):
        super().__init__(0, (target,), (float(variant),), "f64")


class NoisyProduct(DeviceCost):
    """ref test/runtests.jl:105-112: sim((n, du)) = (n*n + du) * (n + randn()*noise); cost = |sim - target|."""
    kind, ndim = K.MODEL_DETERMINISTIC, 2

    def __init__(self, target=5.5, noise=0.01):
        super().__init__(0, (target,), (2.0, float(noise)), "f64")


class Socks(DeviceCost):
    """"Tiny Data, ABC and the Socks of Karl Broman", ref test/runtests.jl:34-44 and :56: theta = (n_socks, prop_pairs);
    n_picked socks are drawn without replacement; cost = |pairs - target[0]| + |odds - target[1]|."""
    kind, ndim = K.MODEL_SOCKS, 2

    def __init__(self, target=(0, 11), n_picked=11):
        super().__init__(0, tuple(target), (float(n_picked),), "f64")


# --------------------------------------------------------------------------------------- results


class Particles:
    """Minimal stand-in for MonteCarloMeasurements.Particles (the reference's result container)."""

    def __init__(self, particles):
        self.particles = np.asarray(particles, dtype=np.float64)

    def __len__(self):
        return self.particles.size

    def mean(self):
        return float(self.particles.mean())

    def std(self):
        return float(self.particles.std(ddof=1)) if self.particles.size > 1 else 0.0

    def approx(self, x: float) -> bool:
        """`p ≈ x` of MonteCarloMeasurements: |mean(p) - x| / std(p) < 2."""
        return abs(self.mean() - x) / self.std() < 2

    def __repr__(self):
        return f"Particles{{Float64,{len(self)}}}({self.mean():.4g} ± {self.std():.2g})"


@dataclass
class SmcResult:
    P: Union[Particles, List[Particles]]
    C: np.ndarray
    eps: float
    iterations: int = 0
    cost_evals: int = 0
    log: list = field(default_factory=list)

    @property
    def ϵ(self):  # noqa: PLC2401  (the reference's field name)
        return self.eps


# --------------------------------------------------------------------------------------- context


class Context:
    """Device + stream + Philox seed: replaces the `rng` keyword of the reference.

    Multi-rank jobs (one process per GPU): pass rank / world and either `nccl_id` (the context then moves the cudaIpc handles
    of its peer arena through NCCL by itself) or `exchange`, a callable `all_gather(bytes) -> list[bytes]` over the ranks
    (MPI, torch.distributed, ...) together with `arena_bytes` -- this mode also lets several ranks share one GPU, which
    NCCL refuses.  NCCL is never used on the data path: ranks meet in flag barriers through NVLink peer memory."""

    def __init__(self, device: int = 0, seed: int = 0x4B49535341424300, rank: int = 0, world: int = 1,
                 nccl_id: Optional[bytes] = None, exchange=None, arena_bytes: int = 0):
        self.L = K.lib()
        self.h = C.c_void_p()
        if world > 1 and nccl_id is not None:
            K.check(self.L.kabc_ctx_create_dist(device, seed, rank, world, nccl_id, C.byref(self.h)))
        elif world > 1:
            if exchange is None or arena_bytes <= 0:
                raise KissABCError(K.ERR_INVALID_ARG, "a multi-rank context needs nccl_id, or exchange + arena_bytes")
            K.check(self.L.kabc_ctx_create_ranks(device, seed, rank, world, C.byref(self.h)))
            buf = C.create_string_buffer(K.IPC_HANDLE_BYTES)
            K.check(self.L.kabc_ctx_arena_export(self.h, int(arena_bytes), buf))
            handles = exchange(buf.raw)
            if len(handles) != world or any(len(x) != K.IPC_HANDLE_BYTES for x in handles):
                raise KissABCError(K.ERR_INVALID_ARG, "exchange() must return one 64-byte handle per rank, in rank order")
            K.check(self.L.kabc_ctx_arena_attach(self.h, b"".join(handles)))
        else:
            K.check(self.L.kabc_ctx_create(device, seed, C.byref(self.h)))
        self.device, self.seed, self.rank, self.world = device, seed, rank, world

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(K.NCCL_ID_BYTES)
        K.check(K.lib().kabc_nccl_unique_id(buf))
        return buf.raw

    def sm_count(self) -> int:
        v = C.c_int()
        K.check(self.L.kabc_ctx_info(self.h, None, None, None, C.byref(v)))
        return v.value

    def close(self, strict: bool = False):
        """Destroys the context.  While smc / ais handles of the context are alive the library refuses (KABC_ERR_STATE) and the
        context stays valid, so that the handles can still be closed afterwards (strict=True raises instead)."""
        if self.h:
            rc = self.L.kabc_ctx_destroy(self.h)
            if rc == K.KABC_OK:
                self.h = C.c_void_p()
            elif strict or rc != K.ERR_STATE:
                K.check(rc)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- thin wrappers over the stateless entry points
    def prior_logpdf(self, prior, theta):
        prior = _as_factored(prior)
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        d, n = theta.shape
        out = np.empty(n)
        K.check(self.L.kabc_prior_logpdf(self.h, prior._pods(), d, K.dptr(theta), n, K.dptr(out)))
        return out

    def prior_sample(self, prior, n, first_id=0, epoch=0):
        prior = _as_factored(prior)
        out = np.empty((len(prior), n))
        K.check(self.L.kabc_prior_sample(self.h, prior._pods(), len(prior), n, first_id, epoch, K.dptr(out)))
        return out

    def eval_cost(self, cost: DeviceCost, theta, first_id=0, epoch=0, return_events=False):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        d, n = theta.shape
        out = np.empty(n)
        ev = np.zeros(n, dtype=np.int64) if return_events else None
        m = cost._pod()
        K.check(self.L.kabc_eval_cost(self.h, C.byref(m), d, K.dptr(theta), n, first_id, epoch, K.dptr(out),
                                      K.i64ptr(ev) if return_events else None))
        return (out, ev) if return_events else out

    def microbench(self, kind: int):
        rate, ms = C.c_double(), C.c_float()
        K.check(self.L.kabc_microbench(self.h, kind, C.byref(rate), C.byref(ms)))
        return rate.value, ms.value


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# --------------------------------------------------------------------------------------- smc


def smc_config(nparticles=100, alpha=0.95, mcmc_retrys=0, mcmc_tol=0.015, epstol=0.0, r_epstol=None,
               min_r_ess=None, max_stretch=2.0, verbose=False, max_iterations=0) -> K.SmcConfigT:
    """Keyword defaults of ref src/smc.jl:95-105."""
    if r_epstol is None:
        r_epstol = (1 - alpha) ** 1.5 / 50
    if min_r_ess is None:
        min_r_ess = alpha * alpha
    return K.SmcConfigT(int(nparticles), float(alpha), int(mcmc_retrys), float(mcmc_tol), float(epstol),
                        float(r_epstol), float(min_r_ess), float(max_stretch), int(bool(verbose)), int(max_iterations))


class SmcSession:
    """Step-by-step smc with the state resident in HBM (kabc_smc_* entry points)."""

    def __init__(self, ctx: Context, prior, cost: DeviceCost, cfg: K.SmcConfigT):
        self.ctx, self.L = ctx, ctx.L
        self.prior = _as_factored(prior)
        self.d, self.N = len(self.prior), int(cfg.nparticles)
        self.h = C.c_void_p()
        m = cost._pod()
        K.check(self.L.kabc_smc_create(ctx.h, self.prior._pods(), self.d, C.byref(m), C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.kabc_smc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        K.check(self.L.kabc_smc_init(self.h))

    def iterate(self) -> int:
        stop = C.c_int()
        K.check(self.L.kabc_smc_iterate(self.h, C.byref(stop)))
        return stop.value

    def iterate_n(self, n: int, ignore_stop: bool = False):
        done, ms = C.c_int(), C.c_float()
        K.check(self.L.kabc_smc_iterate_n(self.h, n, int(ignore_stop), C.byref(done), C.byref(ms)))
        return done.value, ms.value

    def state(self):
        th = np.empty((self.d, self.N)); X = np.empty(self.N); lpi = np.empty(self.N)
        alive = np.empty(self.N, dtype=np.uint8)
        K.check(self.L.kabc_smc_get_state(self.h, K.dptr(th), K.dptr(X), K.dptr(lpi), K.u8ptr(alive)))
        return th, X, lpi, alive

    def set_state(self, th, X, lpi, alive):
        th = np.ascontiguousarray(th, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
        lpi = np.ascontiguousarray(lpi, dtype=np.float64); alive = np.ascontiguousarray(alive, dtype=np.uint8)
        K.check(self.L.kabc_smc_set_state(self.h, K.dptr(th), K.dptr(X), K.dptr(lpi), K.u8ptr(alive)))

    def scalars(self) -> dict:
        eps, flag = C.c_double(), C.c_int32()
        v = [C.c_int64() for _ in range(6)]
        K.check(self.L.kabc_smc_get_scalars(self.h, C.byref(eps), C.byref(flag), *[C.byref(x) for x in v]))
        return dict(eps=eps.value, flag=flag.value, iteration=v[0].value, n_alive=v[1].value, accepted=v[2].value,
                    cost_evals=v[3].value, next_epoch=v[4].value, events=v[5].value)

    def log(self) -> list:
        n = self.L.kabc_smc_get_log(self.h, None, 0)
        if n < 0:
            raise KissABCError(K.ERR_CUDA, "log copy failed")
        buf = (K.SmcLogT * max(n, 1))()
        self.L.kabc_smc_get_log(self.h, buf, n)
        return [dict((f, getattr(buf[i], f)) for f, _ in K.SmcLogT._fields_) for i in range(n)]

    def kernel_launches(self) -> int:
        return int(self.L.kabc_smc_kernel_launches(self.h))

    def bench_steps(self, n: int, flush_bytes: int = 256 << 20) -> np.ndarray:
        """n iterations enqueued back to back, each after an L2 flush outside its timed region: device ms per iteration"""
        buf = (C.c_float * max(n, 1))()
        K.check(self.L.kabc_smc_bench_steps(self.h, n, int(flush_bytes), buf))
        return np.array(buf[:n], dtype=np.float64)

    def profile_iteration(self) -> dict:
        """one iteration with CUDA events between its kernels: warm per-kernel microseconds"""
        buf = (C.c_float * 16)()
        n = C.c_int()
        K.check(self.L.kabc_smc_profile_iteration(self.h, buf, 16, C.byref(n)))
        names = ["sel_1", "sel_2", "cut", "compact_table", "sweep"]
        if n.value == 6:  # two-kernel sweep (Lotka-Volterra, g-and-k, KABC_SWEEP=split)
            names = names[:4] + ["propose", "simulate"]
        return {names[i]: float(buf[i]) for i in range(n.value)}

    def trace_enable(self, on=True):
        K.check(self.L.kabc_smc_trace_enable(self.h, int(on)))

    def trace(self) -> dict:
        N, d = self.N, self.d
        a = np.empty(N, np.int64); b = np.empty(N, np.int64)
        z = np.empty(N); lprob = np.empty(N); lpip = np.empty(N); xp = np.empty(N)
        dec = np.empty(N, np.uint8); thp = np.empty((d, N))
        K.check(self.L.kabc_smc_get_trace(self.h, K.i64ptr(a), K.i64ptr(b), K.dptr(z), K.dptr(lprob), K.dptr(lpip),
                                          K.dptr(xp), K.u8ptr(dec), K.dptr(thp)))
        return dict(a=a, b=b, z=z, lprob=lprob, lpi_p=lpip, xp=xp, decision=dec, theta_p=thp)


def _bundle(theta_rows: np.ndarray):
    """ref src/smc.jl:202-204 / src/KissABC.jl:90-93: one Particles per parameter, scalar if there is only one."""
    P = [Particles(theta_rows[k]) for k in range(theta_rows.shape[0])]
    return P[0] if len(P) == 1 else P


def smc(prior, cost: DeviceCost, *, nparticles=100, alpha=0.95, mcmc_retrys=0, mcmc_tol=0.015, epstol=0.0,
        r_epstol=None, min_r_ess=None, max_stretch=2.0, verbose=False, parallel=False, max_iterations=0,
        ctx: Optional[Context] = None, gather: str = "all") -> SmcResult:
    """smc(prior, cost; kw...) -> (P, C, eps), ref src/smc.jl:92-206.  `parallel` is accepted and ignored
    (the device path is always parallel); `rng` is replaced by the context seed.
    Multi-rank contexts: the call is collective.  gather="all" (default): every rank receives the whole result;
    gather="root": only rank 0 copies the particles out (P and C are None on the other ranks; eps, iterations, log everywhere)."""
    del parallel
    if gather not in ("all", "root"):
        raise KissABCError(K.ERR_INVALID_ARG, "gather must be 'all' or 'root'")
    if not isinstance(cost, DeviceCost):
        raise KissABCError(K.ERR_INVALID_ARG, "the device path needs a registered DeviceCost, not a closure")
    ctx = ctx or default_context()
    prior = _as_factored(prior)
    cfg = smc_config(nparticles, alpha, mcmc_retrys, mcmc_tol, epstol, r_epstol, min_r_ess, max_stretch, verbose,
                     max_iterations)
    d, N = len(prior), int(nparticles)
    want = gather == "all" or ctx.rank == 0
    th = np.empty((d, max(N, 1))) if want else None
    alive = np.empty(max(N, 1), dtype=np.uint8) if want else None
    X = np.empty(max(N, 1)) if want else None
    eps, it, evals = C.c_double(), C.c_int64(), C.c_int64()
    cap = 1 << 13
    logbuf = (K.SmcLogT * cap)()
    m = cost._pod()
    K.check(ctx.L.kabc_smc_run(ctx.h, prior._pods(), d, C.byref(m), C.byref(cfg), K.dptr(th) if want else None,
                               K.u8ptr(alive) if want else None, K.dptr(X) if want else None, C.byref(eps), C.byref(it),
                               C.byref(evals), logbuf, cap))
    n = min(it.value, cap)
    log = [dict((f, getattr(logbuf[i], f)) for f, _ in K.SmcLogT._fields_) for i in range(n)]
    if verbose:
        for r in log:
            print(f"(iteration, ϵ, ESS) = ({r['iteration']}, {r['eps']!r}, {r['n_alive']})")
    if not want:
        return SmcResult(P=None, C=None, eps=eps.value, iterations=it.value, cost_evals=evals.value, log=log)
    mask = alive.astype(bool)
    rows = th if mask.all() else th[:, mask]  # after a resampling iteration everything is alive: no copy of the 8 d N bytes
    return SmcResult(P=_bundle(rows), C=X, eps=eps.value, iterations=it.value, cost_evals=evals.value, log=log)


# --------------------------------------------------------------------------------------- ABCDE, pfilter


@dataclass
class AbcdeResult:
    P: Union[Particles, List[Particles]]
    C: Particles
    reached_eps: bool
    nsim: int = 0
    generations: int = 0

    @property
    def reached_ϵ(self):  # noqa: PLC2401  (the reference's field name)
        return self.reached_eps


def ABCDE(prior, cost: DeviceCost, eps_target, *, nparticles=50, generations=20, alpha=0.0, parallel=False, earlystop=False,
          verbose=False, proposal_width=1.0, ctx: Optional[Context] = None) -> AbcdeResult:
    """ABCDE(prior, cost, ϵ_target; nparticles=50, generations=20, α=0, earlystop=false, proposal_width=1.0)
    -> (P, C, reached_ϵ), ref src/smc.jl:352-428.  `parallel` is accepted and ignored; `rng` is the context seed."""
    del parallel
    if not isinstance(cost, DeviceCost):
        raise KissABCError(K.ERR_INVALID_ARG, "the device path needs a registered DeviceCost, not a closure")
    ctx = ctx or default_context()
    prior = _as_factored(prior)
    d, N = len(prior), int(nparticles)
    cfg = K.AbcdeConfigT(N, int(generations), float(eps_target), float(alpha), float(proposal_width), int(bool(earlystop)), 0)
    th, X = np.empty((d, max(N, 1))), np.empty(max(N, 1))
    reached, nsim, gens = C.c_int32(), C.c_int64(), C.c_int64()
    m = cost._pod()
    K.check(ctx.L.kabc_abcde_run(ctx.h, prior._pods(), d, C.byref(m), C.byref(cfg), K.dptr(th), K.dptr(X), C.byref(reached),
                                 C.byref(nsim), C.byref(gens)))
    if verbose:
        print(f"End: converged = {bool(reached.value)} nsim = {nsim.value} range_ϵ = ({X.min()!r}, {X.max()!r})")
    return AbcdeResult(P=_bundle(th), C=Particles(X), reached_eps=bool(reached.value), nsim=nsim.value, generations=gens.value)


@dataclass
class PfilterResult:
    P: Union[Particles, List[Particles]]
    C: Particles
    eps: float = 0.0
    iterations: int = 0
    nreps: int = 0
    cost_evals: int = 0


def pfilter(prior, cost: DeviceCost, N, *, q=0.7, eff_tol=0.1, epstol=-math.inf, max_iters=math.inf, proposal_width=0.75,
            verbose=False, parallel=False, ctx: Optional[Context] = None) -> PfilterResult:
    """pfilter(prior, cost, N; q=0.7, eff_tol=0.1, epstol=-Inf, max_iters=Inf, proposal_width=0.75) -> (P, C),
    ref src/smc.jl:275-345.  The particle count is raised to ceil((4d+1)/q) when N*q <= 4d, as in the reference."""
    del parallel
    if not isinstance(cost, DeviceCost):
        raise KissABCError(K.ERR_INVALID_ARG, "the device path needs a registered DeviceCost, not a closure")
    ctx = ctx or default_context()
    prior = _as_factored(prior)
    d = len(prior)
    if not (0 < q <= 1):
        raise KissABCError(K.ERR_INVALID_ARG, "pfilter needs 0 < q <= 1")
    n = int(ctx.L.kabc_pfilter_nparticles(int(N), d, float(q)))
    cfg = K.PfilterConfigT(int(N), float(q), float(eff_tol), float(epstol), float(proposal_width),
                           0 if math.isinf(max_iters) else int(max_iters))
    th, X = np.empty((d, n)), np.empty(n)
    eps, it, reps, evals = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64()
    m = cost._pod()
    K.check(ctx.L.kabc_pfilter_run(ctx.h, prior._pods(), d, C.byref(m), C.byref(cfg), K.dptr(th), K.dptr(X), C.byref(eps),
                                   C.byref(it), C.byref(reps), C.byref(evals)))
    if verbose:
        print(f"(iters, ϵ, nreps) = ({it.value}, {eps.value!r}, {reps.value})")
    return PfilterResult(P=_bundle(th), C=Particles(X), eps=eps.value, iterations=it.value, nreps=reps.value,
                         cost_evals=evals.value)


# --------------------------------------------------------------------------------------- AIS


@dataclass(frozen=True)
class AIS:
    """AIS(nparticles), ref src/KissABC.jl:21-23."""
    nparticles: int


@dataclass(frozen=True)
class ApproxKernelizedPosterior:
    """ApproxKernelizedPosterior(prior, cost, target_average_cost), ref src/types.jl:40-49."""
    prior: object
    cost: DeviceCost
    scale: float


@dataclass(frozen=True)
class ApproxPosterior:
    """ApproxPosterior(prior, cost, max_cost), ref src/types.jl:76-104: uniform errors in [-max_cost, max_cost]."""
    prior: object
    cost: DeviceCost
    maxcost: float


def ais_config(nwalkers, nsamples, ntransitions=1, discard_initial=0, thinning=1, retry_sampling=100, scale=1.0,
               posterior=0):
    return K.AisConfigT(int(nwalkers), int(nsamples), int(ntransitions), int(discard_initial), int(thinning),
                        int(retry_sampling), float(scale), int(posterior), 0)


class AisSession:
    """Step-by-step AIS with the ensemble resident in HBM (kabc_ais_* entry points)."""

    def __init__(self, ctx: Context, prior, cost: DeviceCost, cfg: K.AisConfigT):
        self.ctx, self.L = ctx, ctx.L
        self.prior = _as_factored(prior)
        self.d, self.N = len(self.prior), int(cfg.nwalkers)
        self.h = C.c_void_p()
        m = cost._pod()
        K.check(self.L.kabc_ais_create(ctx.h, self.prior._pods(), self.d, C.byref(m), C.byref(cfg), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.kabc_ais_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        K.check(self.L.kabc_ais_init(self.h))

    def sweep(self, n=1) -> float:
        ms = C.c_float()
        K.check(self.L.kabc_ais_sweep(self.h, n, C.byref(ms)))
        return ms.value

    def state(self):
        th = np.empty((self.d, self.N)); lp = np.empty(self.N); ll = np.empty(self.N)
        K.check(self.L.kabc_ais_get_state(self.h, K.dptr(th), K.dptr(lp), K.dptr(ll)))
        return th, lp, ll

    def set_state(self, th, lp, ll):
        th = np.ascontiguousarray(th, dtype=np.float64)
        lp = np.ascontiguousarray(lp, dtype=np.float64); ll = np.ascontiguousarray(ll, dtype=np.float64)
        K.check(self.L.kabc_ais_set_state(self.h, K.dptr(th), K.dptr(lp), K.dptr(ll)))

    def counters(self) -> dict:
        v = [C.c_int64() for _ in range(4)]
        K.check(self.L.kabc_ais_get_counters(self.h, *[C.byref(x) for x in v]))
        return dict(cost_evals=v[0].value, accepted=v[1].value, sweeps=v[2].value, retries=v[3].value)

    def kernel_launches(self) -> int:
        return int(self.L.kabc_ais_kernel_launches(self.h))

    def trace_enable(self, on=True):
        K.check(self.L.kabc_ais_trace_enable(self.h, int(on)))

    def trace(self) -> dict:
        N, d = self.N, self.d
        move = np.empty(N, np.uint8); dec = np.empty(N, np.uint8)
        a = np.empty(N, np.int64); b = np.empty(N, np.int64); c = np.empty(N, np.int64)
        corr = np.empty(N); thp = np.empty((d, N)); lpp = np.empty(N); llp = np.empty(N); e = np.empty(N)
        K.check(self.L.kabc_ais_get_trace(self.h, K.u8ptr(move), K.i64ptr(a), K.i64ptr(b), K.i64ptr(c), K.dptr(corr),
                                          K.dptr(thp), K.dptr(lpp), K.dptr(llp), K.dptr(e), K.u8ptr(dec)))
        return dict(move=move, a=a, b=b, c=c, corr=corr, theta_p=thp, lp_p=lpp, ll_p=llp, e=e, decision=dec)


def sample(model, sampler: AIS, nsamples: int, *, ntransitions=1, discard_initial=0,
           thinning=1, retry_sampling=100, progress=False, ctx: Optional[Context] = None, return_counters=False):
    """sample(model, AIS(N), Ns; ntransitions, discard_initial, thinning, retry_sampling), ref
    src/KissABC.jl:35-94 + AbstractMCMC.sample.  Returns one Particles per parameter (a scalar Particles if d = 1)."""
    del progress
    if not isinstance(model, (ApproxKernelizedPosterior, ApproxPosterior)):
        raise KissABCError(K.ERR_INVALID_ARG, "the device path registers ApproxKernelizedPosterior and ApproxPosterior")
    if not isinstance(model.cost, DeviceCost):
        raise KissABCError(K.ERR_INVALID_ARG, "the device path needs a registered DeviceCost, not a closure")
    ctx = ctx or default_context()
    prior = _as_factored(model.prior)
    d = len(prior)
    hard = isinstance(model, ApproxPosterior)
    cfg = ais_config(sampler.nparticles, nsamples, ntransitions, discard_initial, thinning, retry_sampling,
                     model.maxcost if hard else model.scale, posterior=1 if hard else 0)
    out = np.empty((d, max(int(nsamples), 1)))
    evals, acc = C.c_int64(), C.c_int64()
    m = model.cost._pod()
    K.check(ctx.L.kabc_ais_run(ctx.h, prior._pods(), d, C.byref(m), C.byref(cfg), K.dptr(out), C.byref(evals), C.byref(acc)))
    res = _bundle(out[:, : int(nsamples)])
    if return_counters:
        return res, dict(cost_evals=evals.value, accepted=acc.value)
    return res
