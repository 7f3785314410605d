"""ctypes loader for oracle/libcpu_ref.so (C restatement of the reference CPU path).

TEST INFRASTRUCTURE ONLY -- see oracle/cpu_ref.c header.  Arrays are numpy
uint64 in the C-ABI layout: Fr (n,4), affine (n,8), Jacobian (n,12) or (12,).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcpu_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcpu_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, cols=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def best_multiexp(coeffs, bases, threads: int = 8) -> np.ndarray:
    """arithmetic.rs:465-492 -> Jacobian (12,) u64"""
    coeffs, bases = _c(coeffs, 4), _c(bases, 8)
    assert coeffs.shape[0] == bases.shape[0]
    out = np.zeros(12, dtype=np.uint64)
    rc = lib().ref_best_multiexp(_p(coeffs), _p(bases), ctypes.c_size_t(coeffs.shape[0]), int(threads), _p(out))
    assert rc == 0
    return out


def best_fft(a, omega, log_n: int, threads: int = 8) -> np.ndarray:
    """arithmetic.rs:556-645; returns a transformed copy"""
    a = _c(a, 4).copy()
    assert a.shape[0] == 1 << log_n
    omega = _c(omega).reshape(4)
    rc = lib().ref_best_fft(_p(a), _p(omega), ctypes.c_uint32(log_n), int(threads))
    assert rc == 0
    return a


def ifft(a, omega_inv, divisor, log_n: int, threads: int = 8) -> np.ndarray:
    a = _c(a, 4).copy()
    rc = lib().ref_ifft(_p(a), _p(_c(omega_inv).reshape(4)), _p(_c(divisor).reshape(4)),
                        ctypes.c_uint32(log_n), int(threads))
    assert rc == 0
    return a


def coeff_to_extended(a, k: int, ext_k: int, zeta, zeta_sq, ext_omega, threads: int = 8) -> np.ndarray:
    a = _c(a, 4)
    assert a.shape[0] == 1 << k
    buf = np.zeros((1 << ext_k, 4), dtype=np.uint64)
    buf[: 1 << k] = a
    rc = lib().ref_coeff_to_extended(_p(buf), ctypes.c_uint32(k), ctypes.c_uint32(ext_k),
                                     _p(_c(zeta).reshape(4)), _p(_c(zeta_sq).reshape(4)),
                                     _p(_c(ext_omega).reshape(4)), int(threads))
    assert rc == 0
    return buf


def extended_to_coeff(a, ext_k: int, zeta, zeta_sq, ext_omega_inv, ext_divisor, threads: int = 8) -> np.ndarray:
    a = _c(a, 4).copy()
    assert a.shape[0] == 1 << ext_k
    rc = lib().ref_extended_to_coeff(_p(a), ctypes.c_uint32(ext_k), _p(_c(zeta).reshape(4)),
                                     _p(_c(zeta_sq).reshape(4)), _p(_c(ext_omega_inv).reshape(4)),
                                     _p(_c(ext_divisor).reshape(4)), int(threads))
    assert rc == 0
    return a


def field_vec(field: int, op: int, a, b=None) -> np.ndarray:
    a = _c(a, 4)
    b = a if b is None else _c(b, 4)
    out = np.empty_like(a)
    rc = lib().ref_field_vec(int(field), int(op), _p(a), _p(b), ctypes.c_size_t(a.shape[0]), _p(out))
    assert rc == 0
    return out


def to_mont(field: int, a) -> np.ndarray:
    a = _c(a, 4)
    out = np.empty_like(a)
    lib().ref_to_mont(int(field), _p(a), ctypes.c_size_t(a.shape[0]), _p(out))
    return out


def from_mont(field: int, a) -> np.ndarray:
    a = _c(a, 4)
    out = np.empty_like(a)
    lib().ref_from_mont(int(field), _p(a), ctypes.c_size_t(a.shape[0]), _p(out))
    return out


def g1_mul_gen(k_canonical, threads: int = 8) -> np.ndarray:
    """[k_i] G for canonical 256-bit k_i -> affine (n,8)"""
    k = _c(k_canonical, 4)
    out = np.zeros((k.shape[0], 8), dtype=np.uint64)
    lib().ref_g1_mul_gen(_p(k), ctypes.c_size_t(k.shape[0]), int(threads), _p(out))
    return out


def jac_to_affine(j) -> np.ndarray:
    j = _c(j, 12)
    out = np.zeros((j.shape[0], 8), dtype=np.uint64)
    lib().ref_jac_to_affine(_p(j), ctypes.c_size_t(j.shape[0]), _p(out))
    return out


def msm_naive(coeffs, bases) -> np.ndarray:
    coeffs, bases = _c(coeffs, 4), _c(bases, 8)
    out = np.zeros(12, dtype=np.uint64)
    lib().ref_msm_naive(_p(coeffs), _p(bases), ctypes.c_size_t(coeffs.shape[0]), _p(out))
    return out


def jac_sum(j) -> np.ndarray:
    j = _c(j, 12)
    out = np.zeros(12, dtype=np.uint64)
    lib().ref_jac_sum(_p(j), ctypes.c_size_t(j.shape[0]), _p(out))
    return out


# ---------------------------------------------------------------- input generation
def random_fr_mont(n: int, seed: int) -> np.ndarray:
    """SURVEY 8d generator, vectorised: splitmix64 -> 4 limbs -> mod r -> Montgomery.

    Identical stream to oracle.bn254.random_fr (tested)."""
    idx = np.arange(1, 4 * n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        st = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = st
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    raw = z.reshape(n, 4)
    # reduce mod r: value < 2^256 < 6r, so to_mont(from_mont(.)) style reduction is not
    # enough on its own; use Montgomery mul by R^2 twice:  x -> x*R2/R = xR (reduced),
    # then from_mont gives x mod r canonical.
    m = to_mont(0, raw)          # (x * R) mod r, fully reduced for any 256-bit x
    return m


def random_fr_small_mont(n: int, seed: int, bits: int) -> np.ndarray:
    """uniform scalars < 2^bits (bits <= 64), Montgomery form"""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        st = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = st
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    raw = np.zeros((n, 4), dtype=np.uint64)
    raw[:, 0] = z >> np.uint64(64 - bits) if bits < 64 else z
    return to_mont(0, raw)


# ---- evaluate_h row loop (plonk/evaluation.rs:846-1001 generalised to one flat Calculation list) ----
_QKIND = {"Constant": 0, "Intermediate": 1, "Fixed": 2, "Advice": 3, "Instance": 4, "Aux": 5, "Challenge": 6, "CosetX": 7}
_QCH = {"Beta": 0, "Gamma": 1, "Theta": 2, "Y": 3}


def _qsrc(s):
    return (_QKIND[s[0]], s[1] if len(s) > 1 else 0, s[2] if len(s) > 2 else 0)


def _qcalc_words(c):
    """(op, a.kind, a.index, a.rot, b.kind, b.index, b.rot, challenge, power): the layout of b2_qcalc"""
    t, z = c[0], (0, 0, 0)
    if t in ("Add", "Sub", "Mul"):
        return ({"Add": 0, "Sub": 1, "Mul": 2}[t],) + _qsrc(c[1]) + _qsrc(c[2]) + (0, 0)
    if t == "Negate":
        return (3,) + _qsrc(c[1]) + z + (0, 0)
    if t == "LcChallenge":
        return (4,) + _qsrc(c[1]) + _qsrc(c[2]) + (_QCH[c[3]], c[4])
    if t == "LcTheta":
        return (5,) + _qsrc(c[1]) + _qsrc(c[2]) + (2, 0)
    if t == "MulChAdd":
        return (5,) + _qsrc(c[1]) + _qsrc(c[2]) + (c[3], 0)
    if t == "AddChallenge":
        return (6,) + _qsrc(c[1]) + z + (_QCH[c[2]], 0)
    if t == "Store":
        return (7,) + _qsrc(c[1]) + z + (0, 0)
    raise ValueError(c)


def quotient_eval(rotations, constants, calcs, result, fixed, advice, instance, aux, challenges, log_rows: int,
                  rot_scale: int, x0=None, step=None, threads: int = 8) -> np.ndarray:
    """Every Calculation for every row, then `result` (reference enums as tuples, see oracle/plonk.py).
    constants / challenges / x0 / step: Montgomery (.., 4) arrays; columns: lists of (rows, 4) arrays."""
    rows = 1 << log_rows
    rot = np.asarray(list(rotations), dtype=np.int32)
    cst = _c(constants, 4) if len(constants) else np.zeros((1, 4), np.uint64)
    words = np.asarray([_qcalc_words(c) for c in calcs], dtype=np.uint32).reshape(-1, 9) if len(calcs) \
        else np.zeros((1, 9), np.uint32)
    res = np.asarray(_qsrc(result), dtype=np.uint32)
    keep = []

    def table(cols):
        arrs = [_c(a, 4) for a in cols]
        for a in arrs:
            assert a.shape[0] == rows
        keep.append(arrs)
        t = (ctypes.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
        keep.append(t)
        return t

    ch = _c(challenges, 4) if len(challenges) else np.zeros((1, 4), np.uint64)
    out = np.zeros((rows, 4), dtype=np.uint64)
    x0a = _c(x0).reshape(4) if x0 is not None else None
    sta = _c(step).reshape(4) if step is not None else None
    rc = lib().ref_quotient_eval(_p(rot), ctypes.c_uint32(len(rot)), _p(cst), _p(words), ctypes.c_uint32(len(calcs)),
                                 _p(res), table(fixed), table(advice), table(instance), table(aux), _p(ch),
                                 _p(x0a) if x0a is not None else None, _p(sta) if sta is not None else None,
                                 ctypes.c_uint32(log_rows), ctypes.c_uint32(rot_scale), _p(out), int(threads))
    assert rc == 0
    return out
