"""Pins the oracle: independent KATs (tests/golden/kat.json), every algebraic identity the
reference's own unit tests check, C restatement == Python big-int restatement, and the frozen
fixtures.  No GPU."""
import json
import os
import random

import numpy as np
import pytest

from oracle import bn254 as o
from oracle import cref

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
FIX = np.load(os.path.join(HERE, "golden", "fixtures.npz"))
H = lambda s: int(s, 16)  # noqa: E731


def test_kat_parameters():
    assert o.R_MOD == H(KAT["r"]) and o.Q_MOD == H(KAT["q"])
    limbs = lambda x: ["%016x" % ((x >> (64 * i)) & (2**64 - 1)) for i in range(4)]  # noqa: E731
    assert limbs(o.MONT_R_FR) == KAT["fr_R"]
    assert limbs(o.MONT_R_FR**2 % o.R_MOD) == KAT["fr_R2"]
    assert limbs(o.MONT_R_FQ) == KAT["fq_R"]
    assert limbs(o.MONT_R_FQ**2 % o.Q_MOD) == KAT["fq_R2"]
    assert o.FR_ROOT_OF_UNITY == H(KAT["root_of_unity_2_28"])
    assert pow(o.FR_ROOT_OF_UNITY, 1 << 28, o.R_MOD) == 1 and pow(o.FR_ROOT_OF_UNITY, 1 << 27, o.R_MOD) != 1
    assert pow(7, 1 << 28, o.R_MOD) == H(KAT["delta"])
    for k, name in ((22, "omega_2_22"), (16, "omega_2_16"), (24, "omega_2_24"), (2, "omega_4")):
        assert pow(o.FR_ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD) == H(KAT[name])
    assert o.FR_ZETA_A == H(KAT["zeta_a"]) and o.FR_ZETA_B == H(KAT["zeta_b"])
    assert pow(o.FR_ZETA_A, 3, o.R_MOD) == 1 and o.FR_ZETA_A != 1


def test_kat_curve():
    G = o.G1_GEN
    assert o.g1_is_on_curve(G)
    assert o.g1_mul(G, 2) == (H(KAT["g2x"]), H(KAT["g2y"]))
    assert o.g1_mul(G, 3) == (H(KAT["g3x"]), H(KAT["g3y"]))
    assert o.g1_mul(G, o.R_MOD - 1) == (1, o.Q_MOD - 2)
    assert o.g1_add(o.g1_mul(G, o.R_MOD - 1), G) is None  # r * G = identity
    bases = [o.g1_mul(G, i) for i in range(1, 5)]
    want = (H(KAT["g30x"]), H(KAT["g30y"]))
    assert o.multiexp_serial([1, 2, 3, 4], bases, None) == want
    assert o.best_multiexp([1, 2, 3, 4], bases, 2) == want
    assert o.small_multiexp([1, 2, 3, 4], bases) == want
    assert o.msm_naive([1, 2, 3, 4], bases) == want


def test_kat_ntt4():
    a = list(KAT["ntt4_in"])
    o.best_fft(a, H(KAT["omega_4"]), 2)
    assert a == [H(x) for x in KAT["ntt4_out"]]


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8])
def test_fft_is_dft_and_invertible(k):
    om = pow(o.FR_ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD)
    x = o.random_fr(1 << k, 100 + k)
    y = list(x)
    o.best_fft(y, om, k)
    if k <= 5:
        assert y == o.dft_naive(x, om)
    z = list(y)
    o.EvaluationDomain.ifft(z, o.fr_inv(om), k, o.fr_inv(1 << k))
    assert z == x


def test_reference_test_commit_lagrange():
    """poly/commitment.rs:480-495: commit(lagrange_to_coeff(a)) == commit_lagrange(a), a[i] = i, K = 6."""
    K = 6
    params = o.Params(K, random.Random(7).randrange(1, o.R_MOD))
    domain = o.EvaluationDomain(1, K)
    a = list(range(1 << K))
    b = domain.lagrange_to_coeff(a)
    assert params.commit_lagrange(a) == params.commit(b)
    # commit_lagrange_with_bound drops zeros but commits to the same point (:199-222)
    assert params.commit_lagrange_with_bound(a, 6) == params.commit_lagrange(a)
    coeffs, c = params.commit_lagrange_and_ifft(a, domain.omega_inv, domain.ifft_divisor)
    assert coeffs == b and c == params.commit_lagrange(a)


def test_reference_test_rotate():
    """poly/domain.rs:550-589"""
    domain = o.EvaluationDomain(1, 3)
    rnd = random.Random(11)
    poly = [rnd.randrange(o.R_MOD) for _ in range(8)]
    rot = lambda p, r: [p[(i + r) % 8] for i in range(8)]  # noqa: E731  Polynomial::rotate
    c, c_next, c_prev = (domain.lagrange_to_coeff(rot(poly, r)) for r in (0, 1, -1))
    x = rnd.randrange(o.R_MOD)
    assert o.eval_polynomial(c, x * domain.omega % o.R_MOD) == o.eval_polynomial(c_next, x)
    assert o.eval_polynomial(c, x * domain.omega_inv % o.R_MOD) == o.eval_polynomial(c_prev, x)


def test_reference_test_l_i_and_interpolate():
    """poly/domain.rs:591-619 and arithmetic.rs:932-950"""
    domain = o.EvaluationDomain(1, 3)
    points = [pow(domain.omega, i, o.R_MOD) for i in range(8)]
    ls = []
    for i in range(8):
        e = [0] * 8
        e[i] = 1
        ls.append(o.lagrange_interpolate(points, e))
    rnd = random.Random(13)
    x = rnd.randrange(o.R_MOD)
    xn = pow(x, 8, o.R_MOD)
    ev = domain.l_i_range(x, xn, range(-7, 8))
    for i in range(8):
        assert o.eval_polynomial(ls[i], x) == ev[7 + i]
        assert o.eval_polynomial(ls[(8 - i) % 8], x) == ev[7 - i]
    pts = [rnd.randrange(o.R_MOD) for _ in range(5)]
    evs = [rnd.randrange(o.R_MOD) for _ in range(5)]
    poly = o.lagrange_interpolate(pts, evs)
    assert [o.eval_polynomial(poly, p) for p in pts] == evs


def test_extended_roundtrip_and_zeta_independence():
    """coeff_to_extended then extended_to_coeff returns the zero-padded input, for either cube root."""
    for zeta in (o.FR_ZETA_A, o.FR_ZETA_B):
        dom = o.EvaluationDomain(5, 4, zeta)
        assert dom.extended_k == 6
        c = o.random_fr(16, 21)
        ext = dom.coeff_to_extended(c)
        back = dom.extended_to_coeff(ext)
        assert back[:16] == c and all(v == 0 for v in back[16:])
        # ext[i] = p(zeta * w_ext^i)
        assert ext[5] == o.eval_polynomial(c, zeta * pow(dom.extended_omega, 5, o.R_MOD) % o.R_MOD)


def test_c_oracle_matches_python_field_and_encoding():
    n = 257
    a, b = o.random_fr(n, 1), o.random_fr(n, 2)
    A, B = o.fr_encode(a), o.fr_encode(b)
    assert o.fr_decode(cref.field_vec(0, 0, A, B)) == [x * y % o.R_MOD for x, y in zip(a, b)]
    assert o.fr_decode(cref.field_vec(0, 1, A, B)) == [(x + y) % o.R_MOD for x, y in zip(a, b)]
    assert o.fr_decode(cref.field_vec(0, 2, A, B)) == [(x - y) % o.R_MOD for x, y in zip(a, b)]
    assert o.fr_decode(cref.random_fr_mont(n, 0xB2000001)) == o.random_fr(n, 0xB2000001)
    # Fq
    Rq = o.MONT_R_FQ
    raw = np.array([o._to_limbs(x % o.Q_MOD) for x in a], dtype=np.uint64)
    m = cref.to_mont(1, raw)
    assert [o._limbs_to_int(r) for r in m] == [x % o.Q_MOD * Rq % o.Q_MOD for x in a]
    assert np.array_equal(cref.from_mont(1, m), raw)


@pytest.mark.parametrize("n,threads", [(1, 8), (3, 8), (31, 4), (33, 8), (300, 8), (1000, 3)])
def test_c_oracle_matches_python_msm(n, threads):
    sc = o.random_fr(n, 40 + n)
    ks = np.array([o._to_limbs(x) for x in o.random_fr(n, 50 + n)], dtype=np.uint64)
    pts = cref.g1_mul_gen(ks)
    pp = o.g1_affine_decode(pts)
    assert pp[0] == o.g1_mul(o.G1_GEN, o.random_fr(n, 50 + n)[0])
    want = o.best_multiexp(sc, pp, threads)
    assert o.g1_jacobian_decode(cref.best_multiexp(o.fr_encode(sc), pts, threads)) == want
    if n <= 300:
        assert o.g1_jacobian_decode(cref.msm_naive(o.fr_encode(sc), pts)) == want


def test_c_oracle_msm_adversarial():
    """complete-addition cases: equal bases, inverse pairs, zero scalars, r-1."""
    G = o.G1_GEN
    P = o.g1_mul(G, 12345)
    bases = [P] * 20 + [o.g1_neg(P)] * 20 + [G, o.g1_neg(G), None]
    sc = [5] * 20 + [5] * 20 + [o.R_MOD - 1, o.R_MOD - 1, 77]
    want = o.msm_naive(sc, bases)
    assert want is None
    got = cref.best_multiexp(o.fr_encode(sc), o.g1_affine_encode(bases), 4)
    assert o.g1_jacobian_decode(got) is None
    sc2 = [0, 1, 2, o.R_MOD - 1, (1 << 253) + 5, 0] + [3] * 40
    bases2 = [P] * 6 + [P] * 40
    want2 = o.g1_mul(P, sum(sc2) % o.R_MOD)
    assert o.g1_jacobian_decode(cref.best_multiexp(o.fr_encode(sc2), o.g1_affine_encode(bases2), 8)) == want2
    assert o.best_multiexp(sc2, bases2, 8) == want2


@pytest.mark.parametrize("k", [1, 4, 9, 13, 16])
def test_c_oracle_matches_python_fft(k):
    om = pow(o.FR_ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD)
    x = cref.random_fr_mont(1 << k, 60 + k)
    y = cref.best_fft(x, o.fr_encode([om])[0], k, 8)
    if k <= 9:
        xs = o.fr_decode(x)
        o.best_fft(xs, om, k)
        assert o.fr_decode(y) == xs
    z = cref.ifft(y, o.fr_encode([o.fr_inv(om)])[0], o.fr_encode([o.fr_inv(1 << k)])[0], k, 8)
    assert np.array_equal(z, x)
    assert np.array_equal(cref.best_fft(x, o.fr_encode([om])[0], k, 1), y)  # thread count does not matter


def test_c_oracle_domain_transforms():
    dom = o.EvaluationDomain(5, 6)
    c = o.random_fr(64, 0xB2000012)
    enc = lambda v: o.fr_encode([v])[0]  # noqa: E731
    ext = cref.coeff_to_extended(o.fr_encode(c), 6, dom.extended_k, enc(dom.g_coset), enc(dom.g_coset_inv),
                                 enc(dom.extended_omega), 4)
    assert o.fr_decode(ext) == dom.coeff_to_extended(c)
    back = cref.extended_to_coeff(ext, dom.extended_k, enc(dom.g_coset), enc(dom.g_coset_inv),
                                  enc(dom.extended_omega_inv), enc(dom.extended_ifft_divisor), 4)
    assert o.fr_decode(back)[: 64 * 4] == dom.extended_to_coeff(dom.coeff_to_extended(c))


def test_frozen_fixtures_match_both_oracles():
    x = o.fr_decode(FIX["ntt_k10_in"])
    om = o.fr_decode(FIX["ntt_k10_omega"])[0]
    y = list(x)
    o.best_fft(y, om, 10)
    assert np.array_equal(o.fr_encode(y), FIX["ntt_k10_out"])
    assert np.array_equal(cref.best_fft(FIX["ntt_k10_in"], FIX["ntt_k10_omega"], 10, 8), FIX["ntt_k10_out"])
    want = o.g1_jacobian_decode(FIX["msm_n512_out"])
    assert o.g1_jacobian_decode(cref.best_multiexp(FIX["msm_n512_scalars"], FIX["msm_n512_bases"], 8)) == want
    assert o.g1_jacobian_decode(cref.best_multiexp(o.fr_encode(list(range(64))), FIX["params_k6_g_lagrange"], 8)) == \
        o.g1_jacobian_decode(FIX["params_k6_commit_lagrange"])


def test_point_encoding_roundtrip_and_params_file():
    """oracle restatement of GroupEncoding / Params::write / Params::read (poly/commitment.rs:241-294; [EXT] convention)"""
    import random
    from oracle import bn254 as o
    rng = random.Random(3)
    pts = [None, o.G1_GEN, o.g1_neg(o.G1_GEN)] + [o.g1_mul(o.G1_GEN, rng.randrange(1, o.R_MOD)) for _ in range(20)]
    for sb in (7, 6):
        for p in pts:
            assert o.g1_from_bytes(o.g1_to_bytes(p, sb), sb) == p
    assert o.g1_to_bytes(o.G1_GEN) == (1).to_bytes(32, "little")                 # y = 2 is even: no flag
    assert o.g1_to_bytes(o.g1_neg(o.G1_GEN))[31] == 0x80 and o.g1_to_bytes(None) == bytes(32)
    ref = o.Params(3, 77)
    k, g, gl, extra = o.Params.read(ref.write(b"xyz"))
    assert (k, g, gl, extra) == (3, ref.g, ref.g_lagrange, b"xyz")


def test_shoup_generator_emulation_and_committed_header():
    """tools/gen_shoup.py: the carry chains of the Shoup constant multiplication, executed limb by limb with PTX
    add.cc / madc semantics, reproduce a*w - q*r with q at most 2 below floor(a*w/r) (no carry lost), and the committed
    csrc/fp_shoup.cuh is exactly what the generator prints."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_shoup", os.path.join(root, "tools", "gen_shoup.py"))
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    g = gs.build()
    c = gs.count(g)
    assert c["mad.lo"] == 92 + 16 and c["mad.hi"] + c["mul.hi"] == 92 + 7
    assert gs.selftest(g, n_random=2000) >= 2000
    with open(os.path.join(root, "halo2_gpu_specific_b200", "csrc", "fp_shoup.cuh")) as f:
        assert f.read() == gs.render(g)


# ---- public third-party vectors for this curve (alt_bn128 = BN254): EIP-196 ecAdd / ecMul, EIP-197 pairing check ----
EIP = json.load(open(os.path.join(HERE, "golden", "eip196_197.json")))


def _eip_g1(h):
    x, y = int(h[:64], 16), int(h[64:128], 16)
    return None if (x, y) == (0, 0) else (x, y)


@pytest.mark.parametrize("v", EIP["ecadd"], ids=lambda v: v["name"])
def test_eip196_ecadd(v):
    a, b = _eip_g1(v["input"][:128]), _eip_g1(v["input"][128:256])
    assert o.g1_is_on_curve(a) and o.g1_is_on_curve(b)
    want = _eip_g1(v["expected"])
    assert o.g1_add(a, b) == want
    # the same sum through every MSM restatement (scalars 1, 1), Python and C
    if a is not None and b is not None:
        assert o.multiexp_serial([1, 1], [a, b], None) == want
        got = cref.best_multiexp(o.fr_encode([1, 1]), o.g1_affine_encode([a, b]), 2)
        assert o.g1_jacobian_decode(got) == want


@pytest.mark.parametrize("v", EIP["ecmul"], ids=lambda v: v["name"])
def test_eip196_ecmul(v):
    p, k = _eip_g1(v["input"][:128]), int(v["input"][128:192], 16)
    assert o.g1_is_on_curve(p)
    want = _eip_g1(v["expected"])
    assert o.g1_mul(p, k % o.R_MOD) == want
    assert o.multiexp_serial([k % o.R_MOD], [p], None) == want
    got = cref.best_multiexp(o.fr_encode([k % o.R_MOD]), o.g1_affine_encode([p]), 1)
    assert o.g1_jacobian_decode(got) == want


@pytest.mark.parametrize("v", EIP["pairing"], ids=lambda v: v["name"])
def test_eip197_pairing(v):
    """pins oracle/pairing.py (what the oracle's Decider::verify rests on) to the public pairing-check vectors"""
    from oracle import pairing as pg
    h, pairs = v["input"], []
    for i in range(0, len(h), 384):
        c = [int(h[i + 64 * j: i + 64 * j + 64], 16) for j in range(6)]
        p = None if (c[0], c[1]) == (0, 0) else (c[0], c[1])
        q = ((c[3], c[2]), (c[5], c[4]))                 # EIP-197 encodes the imaginary part first
        assert o.g1_is_on_curve(p) and pg.g2_is_on_curve(q)
        pairs.append((p, q))
    assert pg.pairing_check(pairs) == v["expected"]
    if v["name"] == "jeff1":
        assert pairs[1][1] == pg.G2_GEN                  # the oracle's G2 generator is the standard one
