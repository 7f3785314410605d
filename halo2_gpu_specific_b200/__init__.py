"""B200-native engine for halo2's polynomial-commitment hot path (BN254 MSM + Fr NTT).

Host-side mirror of the reference interface for this path (same names, argument meaning
and error behaviour as halo2_proofs/src/{arithmetic.rs, poly/domain.rs, poly/commitment.rs})
over the C ABI in include/b2pcs.h.  There is no CPU fallback: importing works anywhere,
but every compute call raises if libb2pcs.so or a CUDA device is missing.
"""
from . import _lib  # noqa: F401
from .arithmetic import (  # noqa: F401
    best_fft,
    best_multiexp,
    best_multiexp_gpu_cond,
    gpu_fft,
    gpu_ifft,
    gpu_multiexp,
    gpu_multiexp_async,
    gpu_multiexp_bound,
    gpu_multiexp_bound_and_fft,
    gpu_multiexp_single_gpu_with_bound,
    small_multiexp,
)
from .commitment import Params, ParamsVerifier  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .evaluation import Evaluator, QuotientProgram  # noqa: F401

__all__ = [
    "best_fft", "best_multiexp", "best_multiexp_gpu_cond", "gpu_fft", "gpu_ifft", "gpu_multiexp", "gpu_multiexp_async",
    "gpu_multiexp_bound", "gpu_multiexp_bound_and_fft", "gpu_multiexp_single_gpu_with_bound",
    "Params", "ParamsVerifier", "EvaluationDomain", "small_multiexp", "Evaluator", "QuotientProgram",
]
