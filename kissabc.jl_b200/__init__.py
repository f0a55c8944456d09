"""kissabc.jl_b200 -- B200 (sm_100a) implementation of the KissABC.jl hot path.

Contents: csrc/ (CUDA kernels + the C ABI of libkissabc_cuda.so), build.py (nvcc recipe),
_capi.py (ctypes binding), api.py (host mirror of the reference's Julia interface), julia/ (ccall shim).
Import as `import kissabc_jl_b200` (shim at the repo root).
"""
from ._capi import KissABCError, LIB_PATH, SYMBOLS, lib  # noqa: F401
from .api import (ABCDE, AIS, AbcdeResult, PfilterResult, pfilter, ApproxKernelizedPosterior, ApproxPosterior, AisSession, Beta, Context, Deterministic, DeviceCost,  # noqa: F401
                  DiscreteUniform, Factored, NegativeBinomial, NoisyProduct, Socks,
                  GandK, LotkaVolterra, MA2, Normal, NormalMeanStd, Particles, SmcResult, SmcSession, Truncated,
                  Uniform, ais_config, default_context, sample, smc, smc_config)
from . import _capi, dist, workloads  # noqa: F401

__all__ = [
    "sample", "AIS", "ABCDE", "pfilter", "ApproxKernelizedPosterior", "ApproxPosterior", "Factored", "smc", "Uniform", "Normal", "Truncated", "Beta", "NegativeBinomial",
    "DiscreteUniform", "Particles",
    "DeviceCost", "NormalMeanStd", "MA2", "GandK", "LotkaVolterra", "Deterministic", "NoisyProduct", "Socks", "Context", "SmcSession",
    "AisSession", "KissABCError", "smc_config", "ais_config",
]
