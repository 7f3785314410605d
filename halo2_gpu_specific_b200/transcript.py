"""Host mirror of the prover's Fiat-Shamir transcript: transcript.rs:14-293 of the reference
(Blake2bWrite + Challenge255).  Hashing 65 bytes per commitment is host work in the reference too; nothing here
touches the device.  Points arrive as the engine returns them (Jacobian with Z = 1 or affine, Fq Montgomery
limbs) and are written as the reference writes them: canonical little-endian coordinates into the hash state,
the 32-byte compressed form into the proof.

The compressed form is GroupEncoding::to_bytes of the pinned pairing crate, which is not in the reference tree
([EXT]): x little-endian with the parity of y in bit `sign_bit` of byte 31, as in `Params::write`
(csrc/encoding.cuh); the bit position is a constructor argument.
"""
from __future__ import annotations

import hashlib
from typing import Optional, Tuple

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error

Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
_RQ_INV = pow((1 << 256) % Q_MOD, -1, Q_MOD)

BLAKE2B_PREFIX_CHALLENGE = b"\x00"     # transcript.rs:14
BLAKE2B_PREFIX_POINT = b"\x01"         # :17
BLAKE2B_PREFIX_SCALAR = b"\x02"        # :20

Point = Optional[Tuple[int, int]]


def _fq_from_mont(limbs) -> int:
    v = sum(int(limbs[i]) << (64 * i) for i in range(4))
    return v * _RQ_INV % Q_MOD


def point_from_engine(p) -> Point:
    """(12,) Jacobian with Z in {0, R} (what the commit entry points return) or (8,) affine -> canonical (x, y);
    None for the identity."""
    a = np.asarray(p, dtype=np.uint64).reshape(-1)
    if a.size == 12:
        if not a[8:12].any():
            return None
        z = _fq_from_mont(a[8:12])
        if z != 1:
            raise B2Error(B2_ERR_ARG, "expected a normalised point (Z = 1)")
    elif a.size != 8:
        raise B2Error(B2_ERR_ARG, "expected 8 (affine) or 12 (Jacobian) limbs")
    if not a[:8].any():
        return None
    return (_fq_from_mont(a[0:4]), _fq_from_mont(a[4:8]))


def g1_to_bytes(p: Point, sign_bit: int = 7) -> bytes:
    if p is None:
        return bytes(32)
    b = bytearray(p[0].to_bytes(32, "little"))
    b[31] |= (p[1] & 1) << sign_bit
    return bytes(b)


class Blake2bWrite:
    """transcript.rs:150-226; scalars are canonical Python ints (or (4,) Montgomery limbs), points as above."""

    def __init__(self, sign_bit: int = 7):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")    # :161-164
        self.sign_bit = sign_bit
        self.writer = bytearray()

    @staticmethod
    def _scalar(s) -> int:
        return s % _fr.R_MOD if isinstance(s, int) else _fr.from_mont(s)

    def squeeze_challenge(self) -> int:
        """:197-202 + Challenge255::new (:266-276): from_bytes_wide of the 64-byte digest"""
        self.state.update(BLAKE2B_PREFIX_CHALLENGE)
        return int.from_bytes(self.state.copy().digest(), "little") % _fr.R_MOD

    squeeze_challenge_scalar = squeeze_challenge

    def common_point(self, point) -> None:
        """:204-216"""
        p = point if (point is None or isinstance(point, tuple)) else point_from_engine(point)
        if p is None:
            raise B2Error(B2_ERR_ARG, "cannot write points at infinity to the transcript")
        self.state.update(BLAKE2B_PREFIX_POINT)
        self.state.update(p[0].to_bytes(32, "little"))
        self.state.update(p[1].to_bytes(32, "little"))

    def common_scalar(self, scalar) -> None:
        """:218-223"""
        self.state.update(BLAKE2B_PREFIX_SCALAR)
        self.state.update(self._scalar(scalar).to_bytes(32, "little"))

    def write_point(self, point) -> None:
        """:180-184"""
        p = point if (point is None or isinstance(point, tuple)) else point_from_engine(point)
        self.common_point(p)
        self.writer += g1_to_bytes(p, self.sign_bit)

    def write_scalar(self, scalar) -> None:
        """:185-189"""
        s = self._scalar(scalar)
        self.common_scalar(s)
        self.writer += s.to_bytes(32, "little")

    def finalize(self) -> bytes:
        """:167-171"""
        return bytes(self.writer)
