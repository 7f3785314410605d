"""GPU parity: 256-bit Montgomery field kernels vs the C oracle (bit-exact), through the C ABI."""
import ctypes

import numpy as np
import pytest

from oracle import bn254 as o
from oracle import cref

pytestmark = pytest.mark.gpu


def _vec(gpu, field, op, a, b):
    out = np.empty_like(a)
    gpu.check(gpu.lib().b2_field_vec(field, op, gpu.ptr(a), gpu.ptr(b), a.shape[0], gpu.ptr(out)))
    return out


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("op", [0, 1, 2, 3])
def test_field_ops_random(gpu, field, op):
    n = 1 << 16
    a = cref.random_fr_mont(n, 0xA0 + field)  # reduced mod r < q: valid Montgomery residues for both fields
    b = cref.random_fr_mont(n, 0xB0 + field)
    got = _vec(gpu, field, op, a, b)
    want = cref.field_vec(field, op, a, b)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_edge_values(gpu, field):
    p = o.Q_MOD if field else o.R_MOD
    vals = [0, 1, 2, p - 1, p - 2, (1 << 253), (1 << 254) - 1 if (1 << 254) - 1 < p else p - 3,
            0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, (1 << 128) - 1, p >> 1, (p >> 1) + 1]
    vals = [v % p for v in vals]
    pairs = [(x, y) for x in vals for y in vals]
    a = np.array([o._to_limbs(x) for x, _ in pairs], dtype=np.uint64)
    b = np.array([o._to_limbs(y) for _, y in pairs], dtype=np.uint64)
    for op in range(4):
        assert np.array_equal(_vec(gpu, field, op, a, b), cref.field_vec(field, op, a, b)), op


def test_shoup_constant_multiplication_equals_montgomery_product(gpu):
    """op 4 = the NTT butterflies' constant multiplication (fp_shoup.cuh): derive (w, floor(w 2^256 / r)) from the
    Montgomery-form operand on the device, multiply, reduce: must be the same residue as the generic product."""
    n = 1 << 16
    a = cref.random_fr_mont(n, 0xC0)
    b = cref.random_fr_mont(n, 0xC1)
    assert np.array_equal(_vec(gpu, 0, 4, a, b), cref.field_vec(0, 0, a, b))
    p = o.R_MOD
    vals = [0, 1, 2, p - 1, p - 2, (1 << 253), p - 3, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF, (1 << 128) - 1, (1 << 224) - 1,
            p >> 1, (p >> 1) + 1, int("ffffffff00000000" * 4, 16) % p, int("00000000ffffffff" * 4, 16) % p]
    pairs = [(x, y) for x in vals for y in vals]
    a = np.array([o._to_limbs(x) for x, _ in pairs], dtype=np.uint64)
    b = np.array([o._to_limbs(y) for _, y in pairs], dtype=np.uint64)
    assert np.array_equal(_vec(gpu, 0, 4, a, b), cref.field_vec(0, 0, a, b))


@pytest.mark.parametrize("field", [0, 1])
def test_fused_two_product_reduction(gpu, field):
    """ops 5 / 6: a*b +- b*b with one Montgomery reduction (fp_mul2_add / fp_mul2_sub, the Y3 of the point additions)."""
    p = o.Q_MOD if field else o.R_MOD
    vals = [0, 1, 2, p - 1, p - 2, (1 << 253), p - 3, 0xFFFFFFFF, (1 << 128) - 1, p >> 1, (p >> 1) + 1]
    pairs = [(x, y) for x in vals for y in vals]
    a = np.concatenate([np.array([o._to_limbs(x) for x, _ in pairs], dtype=np.uint64), cref.random_fr_mont(1 << 14, 0xD0 + field)])
    b = np.concatenate([np.array([o._to_limbs(y) for _, y in pairs], dtype=np.uint64), cref.random_fr_mont(1 << 14, 0xD2 + field)])
    ab, bb = cref.field_vec(field, 0, a, b), cref.field_vec(field, 3, b, b)
    assert np.array_equal(_vec(gpu, field, 5, a, b), cref.field_vec(field, 1, ab, bb))
    assert np.array_equal(_vec(gpu, field, 6, a, b), cref.field_vec(field, 2, ab, bb))


def test_shoup_probe_reports_a_rate(gpu):
    muls = ctypes.c_double()
    gpu.check(gpu.lib().b2_shoup_probe(ctypes.byref(muls)))
    assert muls.value > 1e9


def test_imad_probe_reports_a_rate(gpu):
    macs, muls = ctypes.c_double(), ctypes.c_double()
    gpu.check(gpu.lib().b2_imad_probe(ctypes.byref(macs), ctypes.byref(muls)))
    assert muls.value > 1e9 and macs.value == pytest.approx(muls.value * 128)
