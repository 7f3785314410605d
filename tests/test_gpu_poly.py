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


@pytest.mark.parametrize("m,n", [(1, 5), (2, 1), (3, 257), (17, 4096)])
def test_poly_combine_matches_the_horner_fold(gpu, m, n):
    """gwc/prover.rs:47-56: poly_batch = poly_batch * v + poly, element by element"""
    rng = random.Random(m * 1000 + n)
    polys = [[rng.randrange(R) for _ in range(n)] for _ in range(m)]
    for v in (0, 1, R - 1, rng.randrange(R)):
        want = [0] * n
        for p in polys:
            want = [(w * v + c) % R for w, c in zip(want, p)]
        got = arithmetic.poly_combine([enc(p) for p in polys], enc([v])[0])
        assert np.array_equal(got, enc(want))


def test_poly_combine_resident_then_open(gpu):
    """One opening of the multiopen argument on resident polynomials: fold, evaluate, divide; the witness polynomial
    satisfies w(r) * (r - z) + batch(z) == batch(r)."""
    n, m = 1 << 16, 12
    polys = cref.random_fr_mont(n * m, 0xB20000A3).reshape(m, n, 4)
    buf = DeviceBuffer(n * (m + 2)).upload(polys)
    rng = random.Random(11)
    v, z, r = rng.randrange(R), rng.randrange(R), rng.randrange(R)
    L = _lib.lib()
    ptrs = (ctypes.c_void_p * m)(*[buf.ptr + j * n * 32 for j in range(m)])
    d_batch, d_w = buf.ptr + m * n * 32, buf.ptr + (m + 1) * n * 32
    _lib.check(L.b2_poly_combine_dev(ptrs, m, n, _lib.ptr(enc([v])[0]), ctypes.c_void_p(d_batch), None))
    _lib.check(L.b2_kate_division_dev(ctypes.c_void_p(d_batch), n, _lib.ptr(enc([z])[0]), ctypes.c_void_p(d_w), None))
    batch = buf.download(n, m * n)
    assert np.array_equal(batch, arithmetic.poly_combine(list(polys), enc([v])[0]))
    want0 = 0
    for j in range(m):
        want0 = (want0 * v + o.fr_decode(polys[j][:1])[0]) % R
    assert o.fr_decode(batch[:1])[0] == want0
    w = buf.download(n - 1, (m + 1) * n)
    ev = lambda p, x: o.fr_decode(arithmetic.eval_polynomial(p, enc([x])[0])[None])[0]  # noqa: E731
    assert (ev(w, r) * (r - z) + ev(batch, z)) % R == ev(batch, r)
    buf.free()
