"""GPU parity of the z-column construction (batch inversion, prefix product / sum, and the device flows of
permutation::Argument::commit, logup commit_z and shuffle commit_product) against oracle/plonk.py."""
import random

import numpy as np
import pytest

import plonk_fixture as fxm
from oracle import bn254 as o
from oracle import cref

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import grand_product as gp

pytestmark = pytest.mark.gpu
R = o.R_MOD
enc = o.fr_encode


@pytest.mark.parametrize("n", [1, 2, 33, 4096, 100003])
def test_batch_invert(gpu, n):
    rng = random.Random(n)
    vals = [rng.randrange(R) for _ in range(n)]
    for i in range(0, n, 7):
        vals[i] = 0                                   # ff::BatchInvert leaves zeros alone
    got = gp.batch_invert(enc(vals))
    want = [pow(v, -1, R) if v else 0 for v in vals]
    assert np.array_equal(got, enc(want))


def test_batch_invert_large_roundtrip(gpu):
    a = cref.random_fr_mont(1 << 20, 0xB2000041)
    b = gp.batch_invert(a.copy())
    prod = np.empty_like(a)
    h2._lib.check(h2._lib.lib().b2_field_vec(0, 0, h2._lib.ptr(a), h2._lib.ptr(b), a.shape[0], h2._lib.ptr(prod)))
    one = enc([1])[0]
    nz = a.any(axis=1)
    assert np.array_equal(prod[nz], np.broadcast_to(one, prod[nz].shape))
    assert np.array_equal(gp.batch_invert(b.copy()), a)


@pytest.mark.parametrize("op", ["product", "sum"])
@pytest.mark.parametrize("n", [0, 1, 7, 2048, 2049, 70001])
def test_prefix_scan(gpu, op, n):
    rng = random.Random(n * 2 + (op == "sum"))
    vals = [rng.randrange(R) for _ in range(n)]
    init = rng.randrange(R)
    got = gp.prefix_scan(op, enc(vals) if n else np.zeros((0, 4), np.uint64), init)
    want, acc = [init], init
    for v in vals:
        acc = acc * v % R if op == "product" else (acc + v) % R
        want.append(acc)
    assert np.array_equal(got, enc(want))
    if n > 3:   # truncated output and the default start value (the operator's identity)
        got = gp.prefix_scan(op, enc(vals), None, n_out=n - 2)
        want, acc = [], 1 if op == "product" else 0
        for v in [None] + vals[:n - 3]:
            if v is not None:
                acc = acc * v % R if op == "product" else (acc + v) % R
            want.append(acc)
        assert np.array_equal(got, enc(want))


def test_prefix_product_large_telescopes(gpu):
    """2^22 elements: out[i + 1] * inverse(out[i]) == a[i] on a sample, and the total against a host fold of
    the tile totals is covered by the small cases; here the scan must agree with itself across tile borders."""
    n = 1 << 22
    a = cref.random_fr_mont(n, 0xB2000042)
    z = gp.prefix_scan("product", a, 1)
    idx = np.array([0, 1, 2047, 2048, 2049, 4095, 4096, n // 2, n - 2, n - 1])
    zi = o.fr_decode(z[idx])
    zn = o.fr_decode(z[idx + 1])
    ai = o.fr_decode(a[idx])
    for x, y, v in zip(zi, zn, ai):
        assert x * v % R == y


@pytest.fixture(scope="module")
def fx():
    return fxm.build(k=6, seed=17)


def _cols(fx):
    return [enc(c) for c in fx["advice"]], [enc(c) for c in fx["fixed"]], [enc(c) for c in fx["instance"]]


def test_permutation_commit(gpu, fx):
    cs, n = fx["cs"], fx["n"]
    bf = cs.blinding_factors()
    dom = h2.EvaluationDomain(cs.degree(), fx["k"])
    adv, fixed, inst = _cols(fx)
    want = fx["perm_z"]
    blinds = [enc(z[n - bf:]) for z in want]
    got = gp.permutation_commit(dom, cs.permutation_columns, cs.degree(), bf, [enc(s) for s in fx["sigmas"]], adv, fixed,
                                inst, fx["beta"], fx["gamma"], blinds)
    assert len(got) == len(want) == 2
    for g, w in zip(got, want):
        assert np.array_equal(g, enc(w))


def test_logup_commit_z(gpu, fx):
    cs, n = fx["cs"], fx["n"]
    bf = cs.blinding_factors()
    dom = h2.EvaluationDomain(cs.degree(), fx["k"])
    adv, fixed, inst = _cols(fx)
    for lookup, lk in zip(cs.lookups, fx["lookups_lagrange"]):
        got = gp.logup_commit_z(dom, lookup, bf, adv, fixed, inst, enc(lk["m"]), fx["theta"], fx["beta"])
        assert len(got) == len(lk["z"])
        for g, w in zip(got, lk["z"]):
            assert np.array_equal(g, enc(w[:n - bf]))


def test_shuffle_commit_product(gpu, fx):
    cs, n = fx["cs"], fx["n"]
    bf = cs.blinding_factors()
    dom = h2.EvaluationDomain(cs.degree(), fx["k"])
    adv, fixed, inst = _cols(fx)
    for group, z in zip(cs.shuffles, fx["shuffle_z"]):
        got = gp.shuffle_commit_product(dom, group, bf, adv, fixed, inst, fx["theta"], fx["beta"])
        assert np.array_equal(got, enc(z[:n - bf]))
