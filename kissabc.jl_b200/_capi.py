"""ctypes binding of libkissabc_cuda.so (include/kissabc_cuda.h).

There is NO CPU fallback: if the CUDA library is missing or no GPU is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkissabc_cuda.so")

KABC_OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_NCCL, ERR_RETRY_BUDGET, ERR_DEGENERATE, ERR_STATE, ERR_PEER = 1, 2, 3, 4, 5, 6, 7
PRIOR_UNIFORM, PRIOR_NORMAL, PRIOR_TRUNC_NORMAL, PRIOR_BETA, PRIOR_NEG_BINOMIAL, PRIOR_DISCRETE_UNIFORM = 0, 1, 2, 3, 4, 5
MODEL_NORMAL_MEANSTD, MODEL_MA2_AUTOCOV, MODEL_GK_OCTILE, MODEL_LV_SSA, MODEL_DETERMINISTIC, MODEL_SOCKS = 0, 1, 2, 3, 4, 5
F64, F32_ACC64 = 0, 1
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64
MAX_DIM = 16


class PriorT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("p0", C.c_double), ("p1", C.c_double),
                ("lo", C.c_double), ("hi", C.c_double)]


class ModelT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("precision", C.c_int32), ("n_draws", C.c_int32), ("n_target", C.c_int32),
                ("target", C.c_double * 32), ("param", C.c_double * 8)]


class SmcConfigT(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("alpha", C.c_double), ("mcmc_retrys", C.c_int64),
                ("mcmc_tol", C.c_double), ("epstol", C.c_double), ("r_epstol", C.c_double),
                ("min_r_ess", C.c_double), ("max_stretch", C.c_double), ("verbose", C.c_int32),
                ("max_iterations", C.c_int32)]


class AisConfigT(C.Structure):
    _fields_ = [("nwalkers", C.c_int64), ("nsamples", C.c_int64), ("ntransitions", C.c_int64),
                ("discard_initial", C.c_int64), ("thinning", C.c_int64), ("retry_sampling", C.c_int64),
                ("scale", C.c_double), ("posterior", C.c_int32), ("_pad", C.c_int32)]


class AbcdeConfigT(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("generations", C.c_int64), ("eps_target", C.c_double), ("alpha", C.c_double),
                ("proposal_width", C.c_double), ("earlystop", C.c_int32), ("_pad", C.c_int32)]


class PfilterConfigT(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("q", C.c_double), ("eff_tol", C.c_double), ("epstol", C.c_double),
                ("proposal_width", C.c_double), ("max_iters", C.c_int64)]


class SmcLogT(C.Structure):
    _fields_ = [("iteration", C.c_int64), ("eps", C.c_double), ("n_alive", C.c_int64), ("flag", C.c_int32),
                ("resampled", C.c_int32), ("accepted", C.c_int64), ("cost_evals", C.c_int64),
                ("sweeps", C.c_int64)]


class KissABCError(RuntimeError):
    """Mirror of the reference's `error(...)` (an ErrorException): carries the status code and message."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


# every symbol include/kissabc_cuda.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "kabc_version", "kabc_last_error", "kabc_device_count", "kabc_nccl_unique_id", "kabc_ctx_create",
    "kabc_ctx_create_dist", "kabc_ctx_create_ranks", "kabc_ctx_arena_export", "kabc_ctx_arena_attach", "kabc_smc_arena_bytes",
    "kabc_ais_arena_bytes", "kabc_ctx_destroy", "kabc_ctx_info", "kabc_prior_logpdf", "kabc_prior_sample",
    "kabc_eval_cost", "kabc_eval_cost_device", "kabc_smc_run", "kabc_smc_create", "kabc_smc_destroy",
    "kabc_smc_init", "kabc_smc_iterate", "kabc_smc_iterate_n", "kabc_smc_get_state", "kabc_smc_set_state",
    "kabc_smc_get_scalars", "kabc_smc_get_log", "kabc_smc_kernel_launches", "kabc_smc_bench_steps", "kabc_smc_profile_iteration", "kabc_smc_trace_enable",
    "kabc_smc_get_trace", "kabc_ais_run", "kabc_ais_create", "kabc_ais_destroy", "kabc_ais_init",
    "kabc_ais_sweep", "kabc_ais_get_state", "kabc_ais_set_state", "kabc_ais_get_counters",
    "kabc_ais_kernel_launches", "kabc_ais_trace_enable", "kabc_ais_get_trace", "kabc_abcde_run",
    "kabc_pfilter_nparticles", "kabc_pfilter_run", "kabc_microbench",
]

_lib = None


def lib():
    """Load libkissabc_cuda.so.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KissABCError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python kissabc.jl_b200/build.py` "
                                     "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    dp, u8p, i64p, vp = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.c_void_p
    vpp, fp, ip = C.POINTER(C.c_void_p), C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.kabc_version.restype = C.c_int
    L.kabc_last_error.restype = C.c_char_p
    L.kabc_device_count.argtypes = [ip]
    L.kabc_nccl_unique_id.argtypes = [C.c_char_p]
    L.kabc_ctx_create.argtypes = [C.c_int, C.c_uint64, vpp]
    L.kabc_ctx_create_dist.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_char_p, vpp]
    L.kabc_ctx_create_ranks.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int, vpp]
    L.kabc_ctx_arena_export.argtypes = [vp, C.c_uint64, C.c_char_p]
    L.kabc_ctx_arena_attach.argtypes = [vp, C.c_char_p]
    L.kabc_smc_arena_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.kabc_smc_arena_bytes.restype = C.c_uint64
    L.kabc_ais_arena_bytes.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.kabc_ais_arena_bytes.restype = C.c_uint64
    L.kabc_ctx_destroy.argtypes = [vp]
    L.kabc_ctx_info.argtypes = [vp, ip, ip, ip, ip]
    L.kabc_prior_logpdf.argtypes = [vp, C.POINTER(PriorT), C.c_int, dp, C.c_int64, dp]
    L.kabc_prior_sample.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.c_int64, C.c_uint32, C.c_uint32, dp]
    L.kabc_eval_cost.argtypes = [vp, C.POINTER(ModelT), C.c_int, dp, C.c_int64, C.c_uint32, C.c_uint32, dp, i64p]
    L.kabc_eval_cost_device.argtypes = [vp, C.POINTER(ModelT), C.c_int, vp, C.c_int64, C.c_uint32, C.c_uint32, vp, fp]
    L.kabc_smc_run.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(SmcConfigT), dp, u8p, dp,
                               dp, i64p, i64p, C.POINTER(SmcLogT), C.c_int64]
    L.kabc_smc_create.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(SmcConfigT), vpp]
    L.kabc_smc_destroy.argtypes = [vp]
    L.kabc_smc_init.argtypes = [vp]
    L.kabc_smc_iterate.argtypes = [vp, ip]
    L.kabc_smc_iterate_n.argtypes = [vp, C.c_int, C.c_int, ip, fp]
    L.kabc_smc_get_state.argtypes = [vp, dp, dp, dp, u8p]
    L.kabc_smc_set_state.argtypes = [vp, dp, dp, dp, u8p]
    L.kabc_smc_get_scalars.argtypes = [vp, dp, C.POINTER(C.c_int32), i64p, i64p, i64p, i64p, i64p, i64p]
    L.kabc_smc_get_log.argtypes = [vp, C.POINTER(SmcLogT), C.c_int64]
    L.kabc_smc_get_log.restype = C.c_int64
    L.kabc_smc_kernel_launches.argtypes = [vp]
    L.kabc_smc_kernel_launches.restype = C.c_int64
    L.kabc_smc_profile_iteration.argtypes = [vp, fp, C.c_int, ip]
    L.kabc_smc_bench_steps.argtypes = [vp, C.c_int, C.c_uint64, fp]
    L.kabc_smc_trace_enable.argtypes = [vp, C.c_int]
    L.kabc_smc_get_trace.argtypes = [vp, i64p, i64p, dp, dp, dp, dp, u8p, dp]
    L.kabc_ais_run.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(AisConfigT), dp, i64p, i64p]
    L.kabc_ais_create.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(AisConfigT), vpp]
    L.kabc_ais_destroy.argtypes = [vp]
    L.kabc_ais_init.argtypes = [vp]
    L.kabc_ais_sweep.argtypes = [vp, C.c_int, fp]
    L.kabc_ais_get_state.argtypes = [vp, dp, dp, dp]
    L.kabc_ais_set_state.argtypes = [vp, dp, dp, dp]
    L.kabc_ais_get_counters.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.kabc_ais_kernel_launches.argtypes = [vp]
    L.kabc_ais_kernel_launches.restype = C.c_int64
    L.kabc_ais_trace_enable.argtypes = [vp, C.c_int]
    L.kabc_ais_get_trace.argtypes = [vp, u8p, i64p, i64p, i64p, dp, dp, dp, dp, dp, u8p]
    L.kabc_abcde_run.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(AbcdeConfigT), dp, dp,
                                 C.POINTER(C.c_int32), i64p, i64p]
    L.kabc_pfilter_nparticles.argtypes = [C.c_int64, C.c_int, C.c_double]
    L.kabc_pfilter_nparticles.restype = C.c_int64
    L.kabc_pfilter_run.argtypes = [vp, C.POINTER(PriorT), C.c_int, C.POINTER(ModelT), C.POINTER(PfilterConfigT), dp, dp, dp,
                                   i64p, i64p, i64p]
    L.kabc_microbench.argtypes = [vp, C.c_int, dp, fp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != KABC_OK:
        raise KissABCError(rc, lib().kabc_last_error().decode("utf-8", "replace"))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def i64ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def u8ptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))
