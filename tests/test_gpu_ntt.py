"""GPU parity: NTT / EvaluationDomain transforms vs the oracle, bit-exact, through the C ABI."""
import os

import numpy as np
import pytest

import halo2_gpu_specific_b200 as h2
from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu
FIX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures.npz"))
enc = lambda v: o.fr_encode([v])[0]  # noqa: E731


def _omega(k):
    return pow(o.FR_ROOT_OF_UNITY, 1 << (28 - k), o.R_MOD)


def test_golden_fixtures(gpu):
    a = FIX["ntt_k10_in"].copy()
    h2.best_fft(a, FIX["ntt_k10_omega"], 10)
    assert np.array_equal(a, FIX["ntt_k10_out"])
    dom = h2.EvaluationDomain(5, 10)
    b = FIX["ntt_k10_in"].copy()
    assert np.array_equal(dom.lagrange_to_coeff(b), FIX["intt_k10_out"])
    dom6 = h2.EvaluationDomain(5, 6)
    ext = dom6.coeff_to_extended(FIX["ext_k6_in"])
    assert np.array_equal(ext, FIX["ext_k6_out"])
    assert np.array_equal(dom6.extended_to_coeff(ext), FIX["ext_k6_back"])


def test_kat_ntt4(gpu):
    a = o.fr_encode([1, 2, 3, 4])
    h2.best_fft(a, enc(_omega(2)), 2)
    assert o.fr_decode(a) == [0xa, 0x16789af3a83522eb1969386a2f88c094a419fe246c11f9394, o.R_MOD - 2,
                              0x30644e72e131a02850c6967bfe2f29ab91a061a5812d67470242134d2ee06c69]


@pytest.mark.parametrize("k", list(range(1, 15)) + [16, 17, 18, 20])
def test_best_fft_matches_oracle(gpu, k):
    """every pass structure: single pass (k <= 11, 12), two passes (13..22), ragged digit splits"""
    x = cref.random_fr_mont(1 << k, 0xB2000002 + k)
    om = enc(_omega(k))
    a = x.copy()
    h2.best_fft(a, om, k)
    assert np.array_equal(a, cref.best_fft(x, om, k, 8))


@pytest.mark.parametrize("k", [3, 11, 12, 13, 16, 19])
def test_gpu_ifft_matches_oracle_and_roundtrips(gpu, k):
    x = cref.random_fr_mont(1 << k, 0xC0 + k)
    om, omi, div = _omega(k), o.fr_inv(_omega(k)), o.fr_inv(1 << k)
    a = x.copy()
    h2.gpu_ifft(a, enc(omi), k, enc(div))
    assert np.array_equal(a, cref.ifft(x, enc(omi), enc(div), k, 8))
    h2.best_fft(a, enc(om), k)
    assert np.array_equal(a, x)


def test_edge_inputs(gpu):
    k = 13
    om = enc(_omega(k))
    z = np.zeros((1 << k, 4), dtype=np.uint64)
    h2.best_fft(z, om, k)
    assert not z.any()
    # delta at 0 -> all ones; constant -> n * delta
    d = np.zeros((1 << k, 4), dtype=np.uint64)
    d[0] = enc(1)
    h2.best_fft(d, om, k)
    assert np.array_equal(d, np.tile(enc(1), (1 << k, 1)))
    h2.best_fft(d, om, k)
    want = np.zeros((1 << k, 4), dtype=np.uint64)
    want[0] = enc((1 << k) % o.R_MOD)
    assert np.array_equal(d, want)
    # maximum residues r-1 everywhere
    m = np.tile(enc(o.R_MOD - 1), (1 << k, 1))
    got = m.copy()
    h2.best_fft(got, om, k)
    assert np.array_equal(got, cref.best_fft(m, om, k, 8))


@pytest.mark.parametrize("j,k", [(5, 6), (5, 10), (4, 11), (3, 12), (5, 14), (9, 13), (2, 9)])
def test_domain_transforms_match_oracle(gpu, j, k):
    for zeta in (o.FR_ZETA_A, o.FR_ZETA_B):
        dom = h2.EvaluationDomain(j, k, zeta)
        ref = o.EvaluationDomain(j, k, zeta)
        c = cref.random_fr_mont(1 << k, 0xD0 + k)
        ext = dom.coeff_to_extended(c)
        want = cref.coeff_to_extended(c, k, ref.extended_k, enc(ref.g_coset), enc(ref.g_coset_inv),
                                      enc(ref.extended_omega), 8)
        assert np.array_equal(ext, want)
        back = dom.extended_to_coeff(ext)
        wantb = cref.extended_to_coeff(want, ref.extended_k, enc(ref.g_coset), enc(ref.g_coset_inv),
                                       enc(ref.extended_omega_inv), enc(ref.extended_ifft_divisor), 8)
        nq = (1 << k) * (j - 1)
        assert back.shape[0] == nq
        assert np.array_equal(back, wantb[:nq])
        if j > 1:
            assert np.array_equal(back[: 1 << k], c) and not back[1 << k:].any()


def test_divide_by_vanishing_poly(gpu):
    dom = h2.EvaluationDomain(5, 8)
    ref = o.EvaluationDomain(5, 8)
    a = cref.random_fr_mont(1 << dom.extended_k, 0xE1)
    want = o.fr_encode(ref.divide_by_vanishing_poly(o.fr_decode(a)))
    got = dom.divide_by_vanishing_poly(a.copy())
    assert np.array_equal(got, want)


def test_batched_columns(gpu):
    k, cols = 12, 7
    dom = h2.EvaluationDomain(5, k)
    x = cref.random_fr_mont(cols << k, 0xF1).reshape(cols, 1 << k, 4)
    want = np.stack([cref.ifft(x[c], dom.omega_inv, dom.ifft_divisor, k, 8) for c in range(cols)])
    got = dom.lagrange_to_coeff_batch(x.copy())
    assert np.array_equal(got, want)
    ext = dom.coeff_to_extended(got)
    for c in (0, cols - 1):
        assert np.array_equal(ext[c], cref.coeff_to_extended(got[c], k, dom.extended_k, dom.g_coset, dom.g_coset_inv,
                                                              dom.extended_omega, 8))


def test_full_size_properties_k22(gpu):
    """BASELINE size (2^22): round trip, and a sparse input whose transform is known in closed form."""
    k = 22
    n = 1 << k
    om, omi, div = _omega(k), o.fr_inv(_omega(k)), o.fr_inv(n)
    x = cref.random_fr_mont(n, 0xB2000002)
    a = x.copy()
    h2.best_fft(a, enc(om), k)
    h2.gpu_ifft(a, enc(omi), k, enc(div))
    assert np.array_equal(a, x)
    # sparse: x = c0 * delta_{i0} + c1 * delta_{i1}  ->  X[j] = c0 w^(i0 j) + c1 w^(i1 j)
    i0, i1, c0, c1 = 1, 3 * (1 << 20) + 12345, 0x1234567, o.R_MOD - 5
    s = np.zeros((n, 4), dtype=np.uint64)
    s[i0], s[i1] = enc(c0), enc(c1)
    h2.best_fft(s, enc(om), k)
    rng = np.random.default_rng(5)
    js = [0, 1, n - 1, n // 2] + [int(v) for v in rng.integers(0, n, 200)]
    got = o.fr_decode(s[js])
    want = [(c0 * pow(om, i0 * j, o.R_MOD) + c1 * pow(om, i1 * j, o.R_MOD)) % o.R_MOD for j in js]
    assert got == want


def test_best_fft_and_ifft_bit_exact_k22(gpu):
    """BASELINE size, every output element: best_fft and gpu_ifft at k = 22 against the C restatement
    (arithmetic.rs:546-705, 515-534) on all host cores"""
    k = 22
    n = 1 << k
    cores = os.cpu_count() or 8
    om, omi, div = _omega(k), o.fr_inv(_omega(k)), o.fr_inv(n)
    x = cref.random_fr_mont(n, 0xB2000022)
    a = x.copy()
    h2.best_fft(a, enc(om), k)
    assert np.array_equal(a, cref.best_fft(x, enc(om), k, cores))
    b = x.copy()
    h2.gpu_ifft(b, enc(omi), k, enc(div))
    assert np.array_equal(b, cref.ifft(x, enc(omi), enc(div), k, cores))


def test_coset_extend_bit_exact_k22_to_24(gpu):
    """BASELINE size (the bench runs 22 -> 24): coeff_to_extended and extended_to_coeff of one column, every element
    against the C restatement (poly/domain.rs:270-350), both candidate zetas on the forward direction"""
    k = 22
    n = 1 << k
    cores = os.cpu_count() or 8
    c = cref.random_fr_mont(n, 0xB2000024)
    for zi, zeta in enumerate((o.FR_ZETA_A, o.FR_ZETA_B)):
        dom = h2.EvaluationDomain(5, k, zeta)
        assert dom.extended_k == 24
        ext = dom.coeff_to_extended(c)
        want = cref.coeff_to_extended(c, k, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega, cores)
        assert np.array_equal(ext, want)
        if zi:
            continue
        # an arbitrary (not low-degree) extended column: the truncated inverse transform must match element by element
        y = cref.random_fr_mont(1 << dom.extended_k, 0xB2000025)
        back = dom.extended_to_coeff(y)
        wantb = cref.extended_to_coeff(y, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega_inv,
                                       dom.extended_ifft_divisor, cores)
        assert back.shape[0] == n * 4
        assert np.array_equal(back, wantb[: n * 4])
        del y, back, wantb
        rt = dom.extended_to_coeff(ext)
        assert np.array_equal(rt[:n], c) and not rt[n:].any()


def test_three_pass_sizes(gpu):
    """k = 25 needs three passes; check with the sparse closed form and a round trip"""
    k = 25
    n = 1 << k
    om, omi, div = _omega(k), o.fr_inv(_omega(k)), o.fr_inv(n)
    i0, i1, c0, c1 = 5, (1 << 24) + 777, 99, o.R_MOD - 123456789
    s = np.zeros((n, 4), dtype=np.uint64)
    s[i0], s[i1] = enc(c0), enc(c1)
    orig = s.copy()
    h2.best_fft(s, enc(om), k)
    rng = np.random.default_rng(6)
    js = [0, 1, n - 1] + [int(v) for v in rng.integers(0, n, 100)]
    got = o.fr_decode(s[js])
    want = [(c0 * pow(om, i0 * j, o.R_MOD) + c1 * pow(om, i1 * j, o.R_MOD)) % o.R_MOD for j in js]
    assert got == want
    h2.gpu_ifft(s, enc(omi), k, enc(div))
    assert np.array_equal(s, orig)


def test_mixed_host_device_locations(gpu):
    """b2_ntt_exec locations 2 (host in, device out) and 3 (device in, host out)"""
    import ctypes
    from halo2_gpu_specific_b200._lib import NttDesc
    L = gpu.lib()
    k = 13
    n = 1 << k
    dom = h2.EvaluationDomain(5, k)
    x = cref.random_fr_mont(n, 0x77)
    want = cref.coeff_to_extended(x, k, dom.extended_k, dom.g_coset, dom.g_coset_inv, dom.extended_omega, 8)
    d_ext = ctypes.c_void_p()
    gpu.check(L.b2_dev_alloc(dom.extended_len() * 32, ctypes.byref(d_ext)))
    z = np.concatenate([dom.g_coset, dom.g_coset_inv])
    e = NttDesc()
    e.log_n, e.location, e.omega = dom.extended_k, 2, dom.extended_omega.ctypes.data
    e.coset_in = z.ctypes.data
    e.n_in, e.in_stride = n, n
    e.n_out = e.out_stride = dom.extended_len()
    e.columns, e.in_, e.out = 1, x.ctypes.data, d_ext.value
    gpu.check(L.b2_ntt_exec(ctypes.byref(e)))
    got = np.empty_like(want)
    gpu.check(L.b2_memcpy_d2h(gpu.ptr(got), d_ext, got.nbytes))
    assert np.array_equal(got, want)
    # device in -> host out: extended_to_coeff of the resident extended column
    zi = np.concatenate([dom.g_coset_inv, dom.g_coset])
    f = NttDesc()
    f.log_n, f.location, f.omega = dom.extended_k, 3, dom.extended_omega_inv.ctypes.data
    f.divisor, f.coset_out = dom.extended_ifft_divisor.ctypes.data, zi.ctypes.data
    f.n_in = f.in_stride = dom.extended_len()
    f.n_out = f.out_stride = n * dom.quotient_poly_degree
    back = np.empty((n * dom.quotient_poly_degree, 4), dtype=np.uint64)
    f.columns, f.in_, f.out = 1, d_ext.value, back.ctypes.data
    gpu.check(L.b2_ntt_exec(ctypes.byref(f)))
    assert np.array_equal(back[:n], x) and not back[n:].any()
    L.b2_dev_free(d_ext)


@pytest.mark.parametrize("maxm,k", [(7, 20), (5, 18), (8, 17)])
def test_three_and_four_pass_plans(gpu, maxm, k):
    """k >= 27 needs three passes on the device; force small digits (B2_NTT_MAXM) so that the 3- and 4-pass
    index arithmetic (digit reversal, two-level twiddles on later boundaries) is exercised at test sizes"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r);"
        "import halo2_gpu_specific_b200 as h2; from oracle import cref, bn254 as o;"
        "k = %d; dom = h2.EvaluationDomain(5, k); x = cref.random_fr_mont(1 << k, 11);"
        "a = x.copy(); h2.best_fft(a, dom.omega, k); ok1 = np.array_equal(a, cref.best_fft(x, dom.omega, k, 8));"
        "b = x.copy(); dom.lagrange_to_coeff(b); ok2 = np.array_equal(b, cref.ifft(x, dom.omega_inv, dom.ifft_divisor, k, 8));"
        "c = dom.coeff_to_extended(x[: 1 << (k - 2)]) if False else None;"
        "print('PASSES', ok1, ok2)"
    ) % (root, k)
    env = dict(os.environ, B2_NTT_MAXM=str(maxm), B2_NTT_MAXC="11")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "PASSES True True" in out.stdout, out.stdout + out.stderr
