"""Params::write / Params::read (poly/commitment.rs:241-294) with the point (de)compression on the device,
against the oracle's restatement of the encoding ([EXT]: pasta / pairing_bn256 convention, see oracle/bn254.py)."""
import io
import random

import numpy as np
import pytest

from oracle import bn254 as o

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib
from halo2_gpu_specific_b200._lib import B2Error

pytestmark = pytest.mark.gpu


def _points(n, seed):
    rng = random.Random(seed)
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(1, o.R_MOD)) for _ in range(n)]
    pts[1] = None                      # identity
    pts[2] = o.G1_GEN                  # (1, 2): even y
    pts[3] = o.g1_neg(o.G1_GEN)        # (1, q - 2): odd y
    return pts


@pytest.mark.parametrize("sign_bit", [7, 6])
def test_compress_decompress_match_oracle(gpu, sign_bit):
    pts = _points(300, 5)
    aff = o.g1_affine_encode(pts)
    want = b"".join(o.g1_to_bytes(p, sign_bit) for p in pts)
    got = np.empty(32 * len(pts), dtype=np.uint8)
    _lib.check(_lib.lib().b2_g1_compress(_lib.ptr(aff), len(pts), sign_bit, _lib.ptr(got)))
    assert got.tobytes() == want
    back = np.empty_like(aff)
    _lib.check(_lib.lib().b2_g1_decompress(_lib.ptr(got), len(pts), sign_bit, _lib.ptr(back)))
    assert np.array_equal(back, aff)
    assert [o.g1_from_bytes(want[32 * i: 32 * i + 32], sign_bit) for i in range(8)] == pts[:8]


def test_invalid_encodings_are_rejected(gpu):
    pts = _points(64, 6)
    data = bytearray(b"".join(o.g1_to_bytes(p) for p in pts))
    out = np.empty((64, 8), dtype=np.uint64)
    # x with no square root of x^3 + 3
    x = 5
    while o.fq_sqrt((x ** 3 + 3) % o.Q_MOD) is not None:
        x += 1
    bad = bytearray(data)
    bad[32 * 9: 32 * 10] = x.to_bytes(32, "little")
    buf = np.frombuffer(bytes(bad), dtype=np.uint8)
    with pytest.raises(B2Error) as e:
        _lib.check(_lib.lib().b2_g1_decompress(_lib.ptr(buf), 64, 7, _lib.ptr(out)))
    assert "point 9" in str(e.value)
    with pytest.raises(ValueError):
        o.g1_from_bytes(bytes(bad[32 * 9: 32 * 10]))
    # non-canonical x (x + q still fits in 255 bits)
    bad = bytearray(data)
    bad[32 * 20: 32 * 21] = (pts[20][0] + o.Q_MOD).to_bytes(32, "little")
    buf = np.frombuffer(bytes(bad), dtype=np.uint8)
    with pytest.raises(B2Error) as e:
        _lib.check(_lib.lib().b2_g1_decompress(_lib.ptr(buf), 64, 7, _lib.ptr(out)))
    assert "point 20" in str(e.value)


def test_params_write_read_roundtrip(gpu):
    k = 6
    ref = o.Params(k, 0x1234567)
    want_bytes = ref.write(b"halo2-extra")
    p = h2.Params(k, o.g1_affine_encode(ref.g), o.g1_affine_encode(ref.g_lagrange), precompute=False,
                  additional_data=b"halo2-extra")
    w = io.BytesIO()
    p.write(w)
    assert w.getvalue() == want_bytes
    q = h2.Params.read(io.BytesIO(want_bytes))
    assert q.k == k and q.additional_data == b"halo2-extra"
    assert np.array_equal(q.g.read(), o.g1_affine_encode(ref.g))
    assert np.array_equal(q.g_lagrange.read(), o.g1_affine_encode(ref.g_lagrange))
    # the loaded params commit like the original (test_commit_lagrange, poly/commitment.rs:480-495)
    rng = random.Random(1)
    a = [rng.randrange(o.R_MOD) for _ in range(1 << k)]
    got = q.commit_lagrange(o.fr_encode(a))
    assert o.g1_jacobian_decode(got) == ref.commit_lagrange(a)
    with pytest.raises(B2Error):
        h2.Params.read(io.BytesIO(want_bytes[:100]))
    p.free(); q.free()


def test_decompress_large_roundtrip(gpu):
    """2^18 synthetic SRS points: compress from the resident SRS, register the bytes again, read back"""
    from halo2_gpu_specific_b200.arithmetic import Srs
    s = Srs.synthetic(1 << 18, 0, 0xB2000003)
    data = s.read_compressed()
    t = Srs.register_compressed(data, 1 << 18)
    assert np.array_equal(s.read(), t.read())
    s.free(); t.free()
