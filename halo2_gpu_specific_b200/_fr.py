"""Host-side Fr scalar helpers for the reference-interface mirror (domain constants).

Only scalar bookkeeping lives here (EvaluationDomain::new computes a dozen field
constants on the host in the reference too, poly/domain.rs:44-149); vector work goes
through the C ABI.
"""
from __future__ import annotations

import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
S = 28                       # Fr::S
NUM_BITS = 254               # Fr::NUM_BITS
GENERATOR = 7
ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> S, R_MOD)
ZETA = pow(GENERATOR, (R_MOD - 1) // 3, R_MOD)   # default; EvaluationDomain takes zeta as a parameter
_R = (1 << 256) % R_MOD
_RINV = pow(_R, -1, R_MOD)


def to_mont(x: int) -> np.ndarray:
    v = x % R_MOD * _R % R_MOD
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def from_mont(limbs) -> int:
    limbs = np.asarray(limbs, dtype=np.uint64).reshape(4)
    v = sum(int(limbs[i]) << (64 * i) for i in range(4))
    return v * _RINV % R_MOD


def inv(x: int) -> int:
    return pow(x, -1, R_MOD)
