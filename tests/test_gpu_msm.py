"""GPU parity: MSM / Params commits vs the oracle, bit-exact on the (normalised) point, through the C ABI."""
import os

import numpy as np
import pytest

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200.arithmetic import Srs
from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.npz"))
Q_ONE = np.array(o.fq_encode_one(1), dtype=np.uint64)


def _affine(jac):
    """engine output is normalised: Z = 1 (Montgomery) or the identity (0, 1, 0)"""
    jac = np.asarray(jac).reshape(12)
    if not jac[8:].any():
        return None
    assert np.array_equal(jac[8:], Q_ONE), "result must be normalised to Z = 1"
    return o.g1_affine_decode(jac[:8].reshape(1, 8))[0]


def _bases(n, seed):
    ks = np.array([o._to_limbs(x) for x in o.random_fr(n, seed)], dtype=np.uint64) if n <= 4096 else None
    if ks is None:
        ks = cref.from_mont(0, cref.random_fr_mont(n, seed))
    return cref.g1_mul_gen(ks)


def _want(scalars, bases):
    return o.g1_jacobian_decode(cref.best_multiexp(scalars, bases, 8))


def test_golden_fixture(gpu):
    got = h2.best_multiexp(FIX["msm_n512_scalars"], FIX["msm_n512_bases"])
    assert _affine(got) == o.g1_jacobian_decode(FIX["msm_n512_out"])


def test_kat_30G(gpu):
    bases = o.g1_affine_encode([o.g1_mul(o.G1_GEN, i) for i in range(1, 5)])
    got = h2.best_multiexp(o.fr_encode([1, 2, 3, 4]), bases)
    assert _affine(got) == (0x036083bfa420b15a4c11f66a3cffd55318b019feb45f833a876e93848625f5ae,
                            0x2630c348c019c3edb74fe62a7e921361aae9621988223514d56ca8b36adc9e36)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 1000, 4097, (1 << 14) + 1, (1 << 16), (1 << 17) - 1])
def test_msm_matches_oracle_random(gpu, n):
    """n = 2^14 + 1 is the reference's GPU threshold (arithmetic.rs:446); 2^k - 1 is kate_division's length"""
    scalars = cref.random_fr_mont(n, 0xB2000003 + n)
    bases = _bases(n, 0x51 + n)
    assert _affine(h2.best_multiexp(scalars, bases)) == _want(scalars, bases)


@pytest.mark.parametrize("n,c", [(1, 0), (33, 8), (1000, 0), (5000, 13), ((1 << 14) + 1, 0), (1 << 16, 0), (1 << 16, 20)])
def test_msm_window_table_matches_oracle(gpu, n, c):
    """b2_srs_precompute: table of 2^(c*w) multiples, one shared bucket set -- same point"""
    scalars = cref.random_fr_mont(n, 0xB2000003 + n)
    bases = _bases(n, 0x51 + n)
    srs = Srs.register(bases).precompute(c)
    assert _affine(h2.best_multiexp(scalars, srs)) == _want(scalars, bases)
    if n >= 1000:
        assert _affine(h2.best_multiexp(scalars[:777], srs[100:877])) == _want(scalars[:777], bases[100:877])
        small = cref.random_fr_small_mont(n, 9, 16)
        small[::3] = 0
        assert _affine(h2.gpu_multiexp_single_gpu_with_bound(small, srs, 16)) == _want(small, bases)
    srs.free()


def test_msm_window_table_adversarial(gpu):
    G = o.G1_GEN
    P = o.g1_mul(G, 12345)
    bases = [P] * 300 + [o.g1_neg(P)] * 300 + [None, G]
    srs = Srs.register(o.g1_affine_encode(bases)).precompute(9)
    assert _affine(h2.best_multiexp(o.fr_encode([7] * 600 + [5, 0]), srs)) is None
    sc = [o.R_MOD - 1, (1 << 253) - 1, 1, 2] * 150 + [3, 9]
    assert _affine(h2.best_multiexp(o.fr_encode(sc), srs)) == o.msm_naive(sc, bases)
    srs.free()


def test_msm_resident_srs_slices(gpu):
    n = 5000
    scalars = cref.random_fr_mont(n, 3)
    bases = _bases(n, 4)
    srs = Srs.register(bases)
    assert np.array_equal(srs.read(10, 5), bases[10:15])
    assert _affine(h2.best_multiexp(scalars, srs)) == _want(scalars, bases)
    assert _affine(h2.best_multiexp(scalars[:1234], srs[0:1234])) == _want(scalars[:1234], bases[:1234])
    assert _affine(h2.best_multiexp(scalars[:1000], srs[77:1077])) == _want(scalars[:1000], bases[77:1077])
    with pytest.raises(gpu.B2Error):
        h2.best_multiexp(scalars, srs[0:10])
    srs.free()


def test_msm_adversarial_complete_addition(gpu):
    """all-equal bases, (P, -P) with equal digits, zero scalars, r-1, top window all ones, identity base"""
    G = o.G1_GEN
    P = o.g1_mul(G, 12345)
    # 1. everything cancels
    bases = [P] * 300 + [o.g1_neg(P)] * 300
    sc = [7] * 600
    assert _affine(h2.best_multiexp(o.fr_encode(sc), o.g1_affine_encode(bases))) is None
    # 2. all-equal bases, assorted scalars -> [sum] P  (forces P + P inside buckets)
    sc2 = [0, 1, 2, o.R_MOD - 1, (1 << 253) + 5, 0, (1 << 254) - 1 - (1 << 254) % 1] + [3] * 500 + [0xFFFF] * 100
    sc2 = [s % o.R_MOD for s in sc2]
    bases2 = [P] * len(sc2)
    assert _affine(h2.best_multiexp(o.fr_encode(sc2), o.g1_affine_encode(bases2))) == o.g1_mul(P, sum(sc2) % o.R_MOD)
    # 3. identity bases are skipped; zero scalars contribute nothing
    bases3 = [None, G, None, P] * 50
    sc3 = [5, 0, 9, 11] * 50
    assert _affine(h2.best_multiexp(o.fr_encode(sc3), o.g1_affine_encode(bases3))) == o.g1_mul(P, 11 * 50)
    # 4. all-zero scalars
    assert _affine(h2.best_multiexp(o.fr_encode([0] * 100), o.g1_affine_encode([P] * 100))) is None
    # 5. scalars whose every window is 2^c - 1 (carry ripples through all signed windows)
    sc5 = [o.R_MOD - 1, (1 << 253) - 1, (1 << 240) - 1, (1 << 16) - 1, (1 << 15), (1 << 15) - 1] * 20
    pts5 = _bases(len(sc5), 99)
    assert _affine(h2.best_multiexp(o.fr_encode(sc5), pts5)) == _want(o.fr_encode(sc5), pts5)


@pytest.mark.parametrize("bits", [1, 8, 15, 16, 17, 32, 64])
def test_msm_with_bound(gpu, bits):
    """gpu_multiexp_single_gpu_with_bound (arithmetic.rs:334-367): max_bits is a contract"""
    n = 3000
    scalars = cref.random_fr_small_mont(n, 0x77 + bits, bits)
    scalars[::3] = 0  # advice columns are mostly small / zero (commitment.rs:204-212 filters zeros)
    bases = _bases(n, 0x78)
    srs = Srs.register(bases)
    want = _want(scalars, bases)
    assert _affine(h2.gpu_multiexp_single_gpu_with_bound(scalars, srs, bits)) == want
    assert _affine(h2.gpu_multiexp_single_gpu_with_bound(scalars, srs, 254)) == want
    srs.free()


def test_msm_bound_violation_is_an_error(gpu):
    n = 100
    scalars = cref.random_fr_mont(n, 5)
    srs = Srs.register(_bases(n, 6))
    with pytest.raises(gpu.B2Error) as e:
        h2.gpu_multiexp_single_gpu_with_bound(scalars, srs, 16)
    assert e.value.code == gpu.B2_ERR_BOUND
    # the failed call must leave the engine usable (no out-of-bounds writes): same for a window table
    pre = Srs.register(_bases(n, 6)).precompute(10)
    for bits in (1, 9, 16, 20, 100):
        with pytest.raises(gpu.B2Error):
            h2.gpu_multiexp_single_gpu_with_bound(scalars, pre, bits)
    ok = cref.random_fr_small_mont(n, 8, 16)
    assert _affine(h2.gpu_multiexp_single_gpu_with_bound(ok, pre, 16)) == _want(ok, _bases(n, 6))
    pre.free()
    srs.free()


def test_skewed_distribution_one_hot_bucket(gpu):
    """every scalar equal: one bucket per window holds all points (chunks cut it many times)"""
    n = 1 << 15
    scalars = np.tile(o.fr_encode([0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF])[0], (n, 1))
    bases = _bases(n, 0x99)
    assert _affine(h2.best_multiexp(scalars, bases)) == _want(scalars, bases)


def test_params_commit_identity_and_variants(gpu):
    """poly/commitment.rs:480-495 test_commit_lagrange on the GPU path, plus the commit variants"""
    K = 6
    params = h2.Params(K, FIX["params_k6_g"], FIX["params_k6_g_lagrange"])
    dom = h2.EvaluationDomain(1, K)
    a = o.fr_encode(list(range(1 << K)))
    c_lag = params.commit_lagrange(a)
    assert _affine(c_lag) == o.g1_jacobian_decode(FIX["params_k6_commit_lagrange"])
    b = dom.lagrange_to_coeff(a.copy())
    assert np.array_equal(params.commit(b), c_lag)
    assert np.array_equal(params.commit_lagrange_with_bound(a, 6), c_lag)
    coeffs, c2 = params.commit_lagrange_and_ifft(a.copy(), dom.omega_inv, dom.ifft_divisor)
    assert np.array_equal(coeffs, b) and np.array_equal(c2, c_lag)
    # shorter polynomial than n (kate_division output has n - 1 coefficients)
    assert _affine(params.commit(b[:-1])) == _want(b[:-1], FIX["params_k6_g"][:-1])
    params.free()


def test_commit_batch(gpu):
    k, cols = 11, 5
    n = 1 << k
    bases = _bases(n, 0x31)
    params = h2.Params(k, bases, bases)
    dom = h2.EvaluationDomain(5, k)
    x = cref.random_fr_mont(cols * n, 0x32).reshape(cols, n, 4)
    want_pts = [_want(x[c], bases) for c in range(cols)]
    out = params.commit_lagrange_batch(x.copy())
    assert [_affine(p) for p in out] == want_pts
    y = x.copy()
    out2 = params.commit_lagrange_batch(y, ifft=(dom.omega_inv, dom.ifft_divisor))
    assert [_affine(p) for p in out2] == want_pts
    for c in range(cols):
        assert np.array_equal(y[c], cref.ifft(x[c], dom.omega_inv, dom.ifft_divisor, k, 8))
    params.free()


def test_g1_sum(gpu):
    pts = [o.g1_mul(o.G1_GEN, s) for s in (5, 7, o.R_MOD - 12, 0)]
    enc = np.stack([o.g1_jacobian_encode(p) for p in pts])
    assert _affine(h2.arithmetic.g1_sum(enc)) is None
    assert _affine(h2.arithmetic.g1_sum(enc[:2])) == o.g1_mul(o.G1_GEN, 12)
    assert _affine(h2.arithmetic.g1_sum(np.stack([enc[0], enc[0]]))) == o.g1_mul(o.G1_GEN, 10)


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synthetic_multipliers(n, first, seed):
    """h_i of b2_srs_synthetic: bases[i] = [h_i] G"""
    with np.errstate(over="ignore"):
        idx = np.arange(first, first + n, dtype=np.uint64)
        h = _splitmix(np.uint64(seed) ^ _splitmix(idx))
    h[h == 0] = 1
    return h


def dot_mod_r(s_can, h):
    """sum_i s_i * h_i as a Python int: 16-bit pieces so every partial dot fits in uint64 (n <= 2^26)"""
    total = 0
    h16 = [((h >> np.uint64(16 * b)) & np.uint64(0xFFFF)) for b in range(4)]
    for limb in range(4):
        col = s_can[:, limb]
        for a in range(4):
            piece = (col >> np.uint64(16 * a)) & np.uint64(0xFFFF)
            if not piece.any():
                continue
            for b in range(4):
                total += int(np.dot(piece, h16[b])) << (64 * limb + 16 * a + 16 * b)
    return total


def test_synthetic_srs_points(gpu):
    srs = Srs.synthetic(64, first_index=1000, seed=0xB2000003)
    h = synthetic_multipliers(64, 1000, 0xB2000003)
    got = o.g1_affine_decode(srs.read())
    for i in (0, 1, 63):
        assert got[i] == o.g1_mul(o.G1_GEN, int(h[i]))
    srs.free()


@pytest.mark.parametrize("logn,bits", [(20, 254), (22, 254), (22, 16), (24, 254)])
def test_full_size_property(gpu, logn, bits):
    """BASELINE sizes: MSM(s, [h_i]G) must equal [sum s_i h_i mod r] G (linearity; O(n) host check)"""
    n = 1 << logn
    seed = 0xB2000003
    srs = Srs.synthetic(n, 0, seed)
    if logn >= 22:
        srs.precompute()
    if bits == 254:
        scalars = cref.random_fr_mont(n, seed)
    else:
        scalars = cref.random_fr_small_mont(n, seed, bits)
        scalars[::2] = 0  # 50 % zeros column
    got = _affine(h2.gpu_multiexp_single_gpu_with_bound(scalars, srs, bits))
    h = synthetic_multipliers(n, 0, seed)
    total = dot_mod_r(cref.from_mont(0, scalars), h)
    want = o.g1_mul(o.G1_GEN, total % o.R_MOD)
    assert got == want
    srs.free()


@pytest.mark.parametrize("bits,table", [(254, True), (254, False), (16, True)])
def test_msm_2_20_bit_exact_vs_oracle(gpu, bits, table):
    """2^20 points, the whole result against the C restatement of best_multiexp (arithmetic.rs:20-108, 465-492) on
    all host cores: uniform 254-bit scalars (window table and plain bases) and a 16-bit column with 50 % zeros
    (commit_lagrange_with_bound, poly/commitment.rs:199-222)"""
    n = 1 << 20
    seed = 0xB2000020
    cores = os.cpu_count() or 8
    srs = Srs.synthetic(n, 0, seed)
    bases = srs.read()
    if table:
        srs.precompute()
    if bits == 254:
        scalars = cref.random_fr_mont(n, seed + 1)
    else:
        scalars = cref.random_fr_small_mont(n, seed + 2, bits)
        scalars[::2] = 0
    got = _affine(h2.gpu_multiexp_single_gpu_with_bound(scalars, srs, bits))
    assert got == o.g1_jacobian_decode(cref.best_multiexp(scalars, bases, cores))
    srs.free()


def test_cpp_host_mirror_selftest(gpu):
    """host/halo2_b200.hpp (C++ mirror of the Rust interface): the reference's test_commit_lagrange,
    commit variants, iFFT consistency and the coset round trip, compiled by build()"""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "halo2_gpu_specific_b200", "host",
                       "host_selftest")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_selftest ok" in out.stdout


def test_concurrent_callers_share_the_device(gpu):
    """the reference calls this path from rayon workers (plonk/prover.rs:293,470,535,561,643): several host
    threads commit / transform at once; every call takes its own lane (stream + workspace)"""
    from concurrent.futures import ThreadPoolExecutor
    k = 12
    n = 1 << k
    bases = _bases(n, 0x41)
    srs = Srs.register(bases).precompute()
    dom = h2.EvaluationDomain(5, k)
    cols = [cref.random_fr_mont(n, 0x500 + i) for i in range(12)]
    want_pts = [_want(c, bases) for c in cols]
    want_ntt = [cref.ifft(c, dom.omega_inv, dom.ifft_divisor, k, 8) for c in cols]

    def work(i):
        gpu.set_device(0)
        p = h2.best_multiexp(cols[i], srs)
        a = cols[i].copy()
        dom.lagrange_to_coeff(a)
        return _affine(p), a

    with ThreadPoolExecutor(6) as ex:
        res = list(ex.map(work, range(len(cols)))) + list(ex.map(work, range(len(cols))))
    for i, (p, a) in enumerate(res):
        assert p == want_pts[i % len(cols)]
        assert np.array_equal(a, want_ntt[i % len(cols)])
    srs.free()


def test_split_path_for_huge_msm(gpu):
    """MSMs beyond the per-launch limit are split by point range and summed on the device
    (same shape as gpu_multiexp_bound, arithmetic.rs:413-440); B2_MSM_MAX_N shrinks the limit"""
    import subprocess
    import sys
    code = (
        "import os, sys, numpy as np; sys.path.insert(0, %r);"
        "import halo2_gpu_specific_b200 as h2; from halo2_gpu_specific_b200.arithmetic import Srs;"
        "from oracle import cref;"
        "n = 5000; sc = cref.random_fr_mont(n, 3); ks = cref.from_mont(0, cref.random_fr_mont(n, 4));"
        "b = cref.g1_mul_gen(ks); want = cref.jac_to_affine(cref.best_multiexp(sc, b, 4))[0];"
        "ok = [np.array_equal(h2.best_multiexp(sc, s)[:8], want) for s in (Srs.register(b), Srs.register(b).precompute())];"
        "print('SPLIT', ok)"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2_MSM_MAX_N="1024")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "SPLIT [True, True]" in out.stdout, out.stdout + out.stderr


def test_small_multiexp_and_params_verifier(gpu):
    """arithmetic.rs:112-132 (benches/arithmetic.rs shape: a few points) and ParamsVerifier::commit_lagrange
    (poly/commitment.rs:384-389)"""
    n = 16
    bases = _bases(n, 0x61)
    sc = cref.random_fr_mont(n, 0x62)
    want = o.small_multiexp(o.fr_decode(sc), o.g1_affine_decode(bases))
    assert _affine(h2.small_multiexp(sc, bases)) == want
    pv = h2.ParamsVerifier(6, FIX["params_k6_g_lagrange"])
    pub = o.fr_encode([5, 0, 7])
    assert _affine(pv.commit_lagrange(pub)) == _want(pub, FIX["params_k6_g_lagrange"][:3])
    pv.free()


def test_public_eip196_vectors_through_the_engine(gpu):
    """the CUDA MSM / point sum against PUBLIC third-party vectors for this curve (EIP-196 ecAdd / ecMul test vectors,
    tests/golden/eip196_197.json): no oracle between the engine and the expected bytes"""
    import json
    V = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eip196_197.json")))

    def pt(h):
        x, y = int(h[:64], 16), int(h[64:128], 16)
        return None if (x, y) == (0, 0) else (x, y)

    for v in V["ecmul"]:
        p, k = pt(v["input"][:128]), int(v["input"][128:192], 16) % o.R_MOD
        assert _affine(h2.best_multiexp(o.fr_encode([k]), o.g1_affine_encode([p]))) == pt(v["expected"]), v["name"]
    for v in V["ecadd"]:
        a, b = pt(v["input"][:128]), pt(v["input"][128:256])
        want = pt(v["expected"])
        assert _affine(h2.best_multiexp(o.fr_encode([1, 1]), o.g1_affine_encode([a, b]))) == want, v["name"]
        enc = np.stack([o.g1_jacobian_encode(a), o.g1_jacobian_encode(b)])
        assert _affine(h2.arithmetic.g1_sum(enc)) == want, v["name"]


def test_async_msm_matches_blocking_and_oracle(gpu):
    """b2_msm_async / b2_msm_wait: several MSMs in flight from one thread (more tickets than lanes are queued by
    waiting in order), results equal to the blocking call and to the oracle; bound violations surface at wait()"""
    from halo2_gpu_specific_b200 import _lib
    n = 1 << 13
    bases = _bases(n, 0x71)
    srs = Srs.register(bases).precompute()
    cols = [cref.random_fr_mont(n, 0x700 + i) for i in range(5)]
    want = [_want(c, bases) for c in cols]
    pending, got = [], []
    for c in cols:
        if len(pending) == 2:                       # two in flight, as bench.py drives it
            got.append(_affine(pending.pop(0).result()))
        pending.append(h2.gpu_multiexp_async(c, srs))
    got += [_affine(f.result()) for f in pending]
    assert got == want
    assert _affine(h2.gpu_multiexp_async(cols[0][:0], srs[:0]).result()) is None
    small = cref.random_fr_small_mont(n, 5, 16)
    assert _affine(h2.gpu_multiexp_async(small, srs, 16).result()) == _want(small, bases)
    with pytest.raises(_lib.B2Error):
        h2.gpu_multiexp_async(cols[0], srs, 16).result()
    # the lane is free again after a failed wait
    assert _affine(h2.gpu_multiexp_async(cols[1], srs).result()) == want[1]
    srs.free()


def test_affine_batch_probe_sums_match_oracle(gpu):
    """the batched-affine addition probe (Montgomery's trick, one inversion per thread; DESIGN.md 4 'measured') computes
    the group law: every sum P + Q it returns equals the oracle's"""
    import ctypes
    from halo2_gpu_specific_b200 import _lib
    n, B, m = 4096, 64, 192
    P, Q, S = (np.zeros((m, 8), dtype=np.uint64) for _ in range(3))
    b_ms, x_ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib().b2_affine_batch_probe(n, B, ctypes.byref(b_ms), ctypes.byref(x_ms), P.ctypes.data,
                                                Q.ctypes.data, S.ctypes.data, m))
    p, q, s = o.g1_affine_decode(P), o.g1_affine_decode(Q), o.g1_affine_decode(S)
    assert all(o.g1_is_on_curve(x) and x is not None for x in p + q)
    assert [o.g1_add(a, b) for a, b in zip(p, q)] == s


def test_caller_stream_msm_sum_and_phase_events(gpu):
    """what bench.py's multi-GPU step does on one rank: b2_msm_dev on a caller's stream, b2_g1_sum_dev on the same stream
    (it takes another lane: caller-stream calls keep to the lane that carries their stream's work), then the phase
    events of THE MSM's lane are read -- and a second caller-stream MSM lands on the same lane as the first"""
    import ctypes
    from halo2_gpu_specific_b200 import _lib
    L = _lib.lib()
    n = 1 << 12
    scalars = cref.random_fr_mont(n, 0xC0FFEE)
    bases = _bases(n, 77)
    srs = Srs.register(bases)
    try:
        want = _want(scalars, bases)
        d_s, d_p, d_o, st = (ctypes.c_void_p() for _ in range(4))
        _lib.check(L.b2_dev_alloc(n * 32, ctypes.byref(d_s)))
        _lib.check(L.b2_dev_alloc(2 * 96, ctypes.byref(d_p)))
        _lib.check(L.b2_dev_alloc(96, ctypes.byref(d_o)))
        _lib.check(L.b2_memcpy_h2d(d_s, ctypes.c_void_p(scalars.ctypes.data), n * 32))
        _lib.check(L.b2_stream_create(ctypes.byref(st)))
        half = n // 2
        for rep in range(2):
            _lib.check(L.b2_msm_dev(srs.handle, 0, d_s, half, 254, d_p, st))
            _lib.check(L.b2_msm_dev(srs.handle, half, ctypes.c_void_p(d_s.value + half * 32), n - half, 254,
                                    ctypes.c_void_p(d_p.value + 96), st))
            _lib.check(L.b2_g1_sum_dev(d_p, 2, d_o, st))
            _lib.check(L.b2_stream_synchronize(st))
            ph = _lib.last_msm_phases()
            assert ph["total"] > 0 and ph["accumulate"] > 0
        out = np.zeros(12, dtype=np.uint64)
        _lib.check(L.b2_memcpy_d2h(ctypes.c_void_p(out.ctypes.data), d_o, 96))
        _lib.check(L.b2_g1_normalize(ctypes.c_void_p(out.ctypes.data), 1))
        assert _affine(out) == want
        _lib.check(L.b2_stream_destroy(st))
        for p in (d_s, d_p, d_o):
            L.b2_dev_free(p)
    finally:
        srs.free()


def test_g1_sum_groups_dev(gpu):
    """b2_g1_sum_groups_dev: the sums of a rank-major block of partials (what ONE all-gather of every rank's partials for
    a block of columns leaves behind), identities and a P + (-P) column included"""
    import ctypes
    from halo2_gpu_specific_b200 import _lib
    L = _lib.lib()
    ranks, groups = 3, 5
    pts = [[o.g1_mul(o.G1_GEN, 7 + 11 * r + 3 * g) for g in range(groups)] for r in range(ranks)]
    pts[1][2] = None                                   # an identity partial
    pts[0][4], pts[1][4], pts[2][4] = o.g1_mul(o.G1_GEN, 5), o.g1_neg(o.g1_mul(o.G1_GEN, 5)), None   # sums to the identity
    flat = np.stack([o.g1_jacobian_encode(pts[r][g]) for r in range(ranks) for g in range(groups)]).astype(np.uint64)
    d_in, d_out = ctypes.c_void_p(), ctypes.c_void_p()
    _lib.check(L.b2_dev_alloc(flat.nbytes, ctypes.byref(d_in)))
    _lib.check(L.b2_dev_alloc(groups * 96, ctypes.byref(d_out)))
    _lib.check(L.b2_memcpy_h2d(d_in, ctypes.c_void_p(flat.ctypes.data), flat.nbytes))
    _lib.check(L.b2_g1_sum_groups_dev(d_in, ranks, groups, d_out, None))
    _lib.check(L.b2_synchronize())
    out = np.zeros((groups, 12), dtype=np.uint64)
    _lib.check(L.b2_memcpy_d2h(ctypes.c_void_p(out.ctypes.data), d_out, out.nbytes))
    _lib.check(L.b2_g1_normalize(ctypes.c_void_p(out.ctypes.data), groups))
    for g in range(groups):
        want = None
        for r in range(ranks):
            want = o.g1_add(want, pts[r][g])
        assert _affine(out[g]) == want
    L.b2_dev_free(d_in)
    L.b2_dev_free(d_out)
