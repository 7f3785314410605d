"""GPU parity of eval_polynomial and kate_division (halo2_proofs/src/arithmetic.rs:707-773) against the oracle."""
import ctypes
import random

import numpy as np
import pytest

from oracle import bn254 as o
from oracle import cref

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib, arithmetic
from halo2_gpu_specific_b200.evaluation import DeviceBuffer

pytestmark = pytest.mark.gpu
R = o.R_MOD
enc = o.fr_encode


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 4097, 1 << 16])
def test_eval_polynomial(gpu, n):
    rng = random.Random(n)
    a = [rng.randrange(R) for _ in range(n)]
    for x in (0, 1, rng.randrange(R)):
        got = arithmetic.eval_polynomial(enc(a), enc([x])[0])
        assert o.fr_decode(got[None])[0] == o.eval_polynomial(a, x)


def test_eval_polynomial_batch_resident(gpu):
    n, cols = 1 << 14, 9
    polys = cref.random_fr_mont(n * cols, 0xB20000A1).reshape(cols, n, 4)
    buf = DeviceBuffer(n * cols).upload(polys)
    x = 0x1234567890ABCDEF1234567890ABCDEF % R
    out = np.empty((cols, 4), dtype=np.uint64)
    _lib.check(_lib.lib().b2_eval_polynomial_dev(ctypes.c_void_p(buf.ptr), cols, n, n, _lib.ptr(enc([x])[0]), _lib.ptr(out)))
    for c in range(cols):
        assert o.fr_decode(out[c][None])[0] == o.eval_polynomial(o.fr_decode(polys[c]), x)
    buf.free()


@pytest.mark.parametrize("n", [2, 3, 64, 2049, 5000])
def test_kate_division(gpu, n):
    rng = random.Random(n)
    a = [rng.randrange(R) for _ in range(n)]
    for b in (0, 1, rng.randrange(R)):
        got = arithmetic.kate_division(enc(a), enc([b])[0])
        assert np.array_equal(got, enc(o.kate_division(a, b)))


def test_kate_division_large_identity(gpu):
    """2^20 coefficients: q(r) * (r - b) + a(b) == a(r) at a random r"""
    n = 1 << 20
    a = cref.random_fr_mont(n, 0xB20000A2)
    rng = random.Random(7)
    b, r = rng.randrange(R), rng.randrange(R)
    q = arithmetic.kate_division(a, enc([b])[0])
    ev = lambda p, x: o.fr_decode(arithmetic.eval_polynomial(p, enc([x])[0])[None])[0]  # noqa: E731
    assert (ev(q, r) * (r - b) + ev(a, b)) % R == ev(a, r)
