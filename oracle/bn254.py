"""CPU oracle for the halo2 polynomial-commitment hot path (BN254 / KZG).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.

PARITY STATUS: **byte-level parity unpinned.**  The reference
(DelphinusLab/halo2-gpu-specific) is Rust; there is no Rust toolchain in the
build image, the field/curve arithmetic lives in an un-vendored git dependency
(``pairing_bn256 0.1.1`` @ lanbones/pairing rev 30b052f2, Cargo.lock:1284-1286)
and the reference's tests hold no golden byte vectors (every test draws from
OsRng and checks algebraic identities).  This oracle therefore restates the
*published* BN254 parameters and the reference's *algorithms*, and is pinned
against (a) independent known-answer vectors (2G, 3G,
(r-1)G, rG, a 4-term MSM, roots of unity, NTT_4([1,2,3,4])), and (b) every
algebraic identity the reference's own unit tests check
(poly/commitment.rs:480-495 test_commit_lagrange, poly/domain.rs:550-619
test_rotate / test_l_i, arithmetic.rs:932-950 test_lagrange_interpolate).
See tests/test_oracle.py.

All functions are plain Python big-int; they are meant for small sizes
(n <= 2^12 or so).  The C restatement in oracle/cpu_ref.c covers larger sizes.

Element encodings (what crosses the C ABI, see include/b2pcs.h):
  Fr / Fq  : 4 x u64 little-endian limbs, Montgomery form a*2^256 mod p
             (reference: transmute::<_, &[u64;4]> in plonk/prover.rs:176).
  G1Affine : x || y (64 B), identity encoded as (0, 0).
  G1       : X || Y || Z Jacobian (96 B), x = X/Z^2, y = Y/Z^3, identity Z = 0.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# Published BN254 (alt_bn128) parameters
# --------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
CURVE_B = 3  # y^2 = x^3 + 3
G1_GEN = (1, 2)
FR_S = 28  # two-adicity of r - 1
FR_GENERATOR = 7  # multiplicative generator used by the ff derive for bn256 Fr
MONT_BITS = 256
MONT_R_FR = (1 << MONT_BITS) % R_MOD
MONT_R_FQ = (1 << MONT_BITS) % Q_MOD
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)  # order 2^28
# The two primitive cube roots of unity in Fr.  Which one the pinned crate
# exports as Fr::ZETA is not recoverable from the reference tree,
# so every API takes zeta as a parameter; ZETA is only the default.
FR_ZETA_A = pow(FR_GENERATOR, (R_MOD - 1) // 3, R_MOD)
FR_ZETA_B = FR_ZETA_A * FR_ZETA_A % R_MOD
FR_ZETA = FR_ZETA_A


def fr_inv(a: int) -> int:
    return pow(a, -1, R_MOD)


def fq_inv(a: int) -> int:
    return pow(a, -1, Q_MOD)


# --------------------------------------------------------------------------
# Encodings (numpy <-> int)
# --------------------------------------------------------------------------
def _to_limbs(x: int) -> List[int]:
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def fr_encode(vals: Sequence[int]) -> np.ndarray:
    """canonical ints -> (n,4) u64 Montgomery limbs"""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = _to_limbs(v % R_MOD * MONT_R_FR % R_MOD)
    return out


_RINV_FR = pow(MONT_R_FR, -1, R_MOD)
_RINV_FQ = pow(MONT_R_FQ, -1, Q_MOD)


def _limbs_to_int(row) -> int:
    return int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)


def fr_decode(arr: np.ndarray) -> List[int]:
    """(n,4) u64 Montgomery limbs -> canonical ints"""
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [_limbs_to_int(r) * _RINV_FR % R_MOD for r in arr]


def fq_encode_one(v: int) -> List[int]:
    return _to_limbs(v % Q_MOD * MONT_R_FQ % Q_MOD)


def g1_affine_encode(pts: Sequence[Optional[Tuple[int, int]]]) -> np.ndarray:
    """affine points (None = identity) -> (n,8) u64: x || y in Fq Montgomery"""
    out = np.zeros((len(pts), 8), dtype=np.uint64)
    for i, p in enumerate(pts):
        if p is None:
            continue
        out[i, :4] = fq_encode_one(p[0])
        out[i, 4:] = fq_encode_one(p[1])
    return out


def g1_affine_decode(arr: np.ndarray) -> List[Optional[Tuple[int, int]]]:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for r in arr:
        x = _limbs_to_int(r[:4]) * _RINV_FQ % Q_MOD
        y = _limbs_to_int(r[4:]) * _RINV_FQ % Q_MOD
        out.append(None if (x == 0 and y == 0) else (x, y))
    return out


def g1_jacobian_decode(arr: np.ndarray) -> Optional[Tuple[int, int]]:
    """96-byte Jacobian (12 u64, Montgomery) -> affine point or None.

    Projective results are not unique; the reference always normalises
    (to_affine / batch_normalize, plonk/prover.rs:130,304,484) before the
    transcript, so parity is defined on the affine point.
    """
    r = np.asarray(arr, dtype=np.uint64).reshape(12)
    X = _limbs_to_int(r[0:4]) * _RINV_FQ % Q_MOD
    Y = _limbs_to_int(r[4:8]) * _RINV_FQ % Q_MOD
    Z = _limbs_to_int(r[8:12]) * _RINV_FQ % Q_MOD
    if Z == 0:
        return None
    zi = fq_inv(Z)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 % Q_MOD * zi % Q_MOD)


def g1_jacobian_encode(p: Optional[Tuple[int, int]]) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    if p is None:
        out[4:8] = fq_encode_one(1)
        return out
    out[0:4] = fq_encode_one(p[0])
    out[4:8] = fq_encode_one(p[1])
    out[8:12] = fq_encode_one(1)
    return out


# --------------------------------------------------------------------------
# Deterministic input generation (shared with the C oracle and the CUDA side)
# --------------------------------------------------------------------------
_M64 = 0xFFFFFFFFFFFFFFFF


def splitmix64(state: int) -> Tuple[int, int]:
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def random_fr(n: int, seed: int) -> List[int]:
    """splitmix64 -> 4 limbs -> reduce mod r (canonical ints)."""
    out = []
    st = seed & _M64
    for _ in range(n):
        v = 0
        for j in range(4):
            st, z = splitmix64(st)
            v |= z << (64 * j)
        out.append(v % R_MOD)
    return out


# --------------------------------------------------------------------------
# G1 arithmetic (affine, complete): the reference uses the complete `Curve`
# operators inside buckets (arithmetic.rs:66-75), so P+P, P+(-P) and identity
# operands are all legal.
# --------------------------------------------------------------------------
Point = Optional[Tuple[int, int]]


def g1_is_on_curve(p: Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(p: Point) -> Point:
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_add(p: Point, q: Point) -> Point:
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * fq_inv(2 * y1) % Q_MOD
    else:
        lam = (y2 - y1) * fq_inv(x2 - x1) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_double(p: Point) -> Point:
    return g1_add(p, p)


def g1_mul(p: Point, k: int) -> Point:
    k %= R_MOD
    acc: Point = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_double(p)
        k >>= 1
    return acc


def msm_naive(scalars: Sequence[int], bases: Sequence[Point]) -> Point:
    acc: Point = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


# --------------------------------------------------------------------------
# MSM: restatement of arithmetic.rs
# --------------------------------------------------------------------------
def _window_bits(m: int) -> int:
    """arithmetic.rs:23-29: c = 1 if m<4; 3 if m<32; else ceil(ln m)."""
    if m < 4:
        return 1
    if m < 32:
        return 3
    return int(math.ceil(math.log(float(m))))


def _get_at(segment: int, c: int, repr_bytes: bytes) -> int:
    """arithmetic.rs:31-49: unsigned c-bit digit through an 8-byte LE window."""
    skip_bits = segment * c
    skip_bytes = skip_bits // 8
    if skip_bytes >= 32:
        return 0
    v = bytearray(8)
    chunk = repr_bytes[skip_bytes:skip_bytes + 8]
    v[: len(chunk)] = chunk
    tmp = int.from_bytes(v, "little")
    tmp >>= skip_bits - skip_bytes * 8
    return tmp % (1 << c)


def multiexp_serial(coeffs: Sequence[int], bases: Sequence[Point], acc: Point) -> Point:
    """arithmetic.rs:20-108 (Pippenger, unsigned digits, MSB-first segments)."""
    reprs = [int(s % R_MOD).to_bytes(32, "little") for s in coeffs]  # to_repr(), :21
    c = _window_bits(len(bases))
    segments = 256 // c + 1  # :51
    for seg in reversed(range(segments)):
        for _ in range(c):
            acc = g1_double(acc)  # :54-56
        buckets: List[Point] = [None] * ((1 << c) - 1)  # :89
        for rp, base in zip(reprs, bases):
            d = _get_at(seg, c, rp)
            if d != 0:
                buckets[d - 1] = g1_add(buckets[d - 1], base)  # :91-96
        running: Point = None
        for b in reversed(buckets):  # :102-106 summation by parts
            running = g1_add(b, running)
            acc = g1_add(acc, running)
    return acc


def best_multiexp(coeffs: Sequence[int], bases: Sequence[Point], num_threads: int = 8) -> Point:
    """arithmetic.rs:465-492: chunk = n / T, ordered fold of per-chunk results."""
    assert len(coeffs) == len(bases)  # :466
    n = len(coeffs)
    if n > num_threads:
        chunk = n // num_threads
        results = []
        for lo in range(0, n, chunk):
            results.append(multiexp_serial(coeffs[lo:lo + chunk], bases[lo:lo + chunk], None))
        acc: Point = None
        for r in results:
            acc = g1_add(acc, r)
        return acc
    return multiexp_serial(coeffs, bases, None)


def small_multiexp(coeffs: Sequence[int], bases: Sequence[Point]) -> Point:
    """arithmetic.rs:112-132: shared-doubling double-and-add."""
    reprs = [int(s % R_MOD).to_bytes(32, "little") for s in coeffs]
    acc: Point = None
    for byte_idx in reversed(range(32)):
        for bit_idx in reversed(range(8)):
            acc = g1_double(acc)
            for i, rp in enumerate(reprs):
                if (rp[byte_idx] >> bit_idx) & 1:
                    acc = g1_add(acc, bases[i])
    return acc


def best_multiexp_gpu_cond(coeffs, bases, num_threads: int = 8) -> Point:
    """arithmetic.rs:442-458 (the dispatcher Params::commit* call)."""
    if len(coeffs) == 0:
        return None
    return best_multiexp(coeffs, bases, num_threads)


# --------------------------------------------------------------------------
# NTT: restatement of best_fft_cpu (arithmetic.rs:556-645)
# --------------------------------------------------------------------------
def _bitreverse(n: int, l: int) -> int:
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def best_fft(a: List[int], omega: int, log_n: int) -> None:
    """In-place radix-2 DIT NTT, natural in -> natural out (arithmetic.rs:556-641).

    Field elements are canonical, so any exact DFT gives the same bits; this
    follows the reference's iterative branch (:613-641)."""
    n = len(a)
    assert n == 1 << log_n  # :569
    for k in range(n):  # :571-576
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    twiddles = [1] * max(n // 2, 1)  # :580-611
    for i in range(1, n // 2):
        twiddles[i] = twiddles[i - 1] * omega % R_MOD
    chunk = 2
    twiddle_chunk = n // 2
    for _ in range(log_n):  # :616-641
        half = chunk // 2
        for base in range(0, n, chunk):
            for i in range(half):
                t = a[base + half + i] * twiddles[i * twiddle_chunk] % R_MOD
                u = a[base + i]
                a[base + i] = (u + t) % R_MOD
                a[base + half + i] = (u - t) % R_MOD
        chunk *= 2
        twiddle_chunk //= 2


def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, R_MOD) for j in range(n)) % R_MOD for i in range(n)]


# --------------------------------------------------------------------------
# EvaluationDomain: restatement of poly/domain.rs
# --------------------------------------------------------------------------
class EvaluationDomain:
    """poly/domain.rs:24-149 (constants) and :233-423 (transforms)."""

    def __init__(self, j: int, k: int, zeta: int = FR_ZETA):
        self.quotient_poly_degree = j - 1  # :46
        self.k = k
        self.n = 1 << k
        extended_k = k
        while (1 << extended_k) < self.n * self.quotient_poly_degree:  # :56-59
            extended_k += 1
        self.extended_k = extended_k
        ext_omega = FR_ROOT_OF_UNITY
        for _ in range(extended_k, FR_S):  # :66-68
            ext_omega = ext_omega * ext_omega % R_MOD
        omega = ext_omega
        for _ in range(k, extended_k):  # :78-80
            omega = omega * omega % R_MOD
        self.omega = omega
        self.omega_inv = fr_inv(omega)
        self.extended_omega = ext_omega
        self.extended_omega_inv = fr_inv(ext_omega)
        self.g_coset = zeta  # :88
        self.g_coset_inv = zeta * zeta % R_MOD  # :89
        # t_evaluations :91-114, inverted :124-131
        orig = pow(zeta, self.n, R_MOD)
        step = pow(ext_omega, self.n, R_MOD)
        t_ev = []
        cur = orig
        while True:
            t_ev.append(cur)
            cur = cur * step % R_MOD
            if cur == orig:
                break
        assert len(t_ev) == 1 << (extended_k - k)  # :105
        self.t_evaluations = [fr_inv((t - 1) % R_MOD) for t in t_ev]
        self.ifft_divisor = fr_inv((1 << k) % R_MOD)  # :116
        self.extended_ifft_divisor = fr_inv((1 << extended_k) % R_MOD)  # :117
        self.barycentric_weight = fr_inv(self.n % R_MOD)  # :121

    def extended_len(self) -> int:
        return 1 << self.extended_k

    @staticmethod
    def ifft(a: List[int], omega_inv: int, log_n: int, divisor: int) -> None:
        """:400-414"""
        best_fft(a, omega_inv, log_n)
        for i in range(len(a)):
            a[i] = a[i] * divisor % R_MOD

    def lagrange_to_coeff(self, a: Sequence[int]) -> List[int]:
        """:233-243"""
        a = list(a)
        assert len(a) == 1 << self.k
        self.ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    def distribute_powers_zeta(self, a: List[int], into_coset: bool) -> None:
        """:382-398"""
        powers = [self.g_coset, self.g_coset_inv] if into_coset else [self.g_coset_inv, self.g_coset]
        for idx in range(len(a)):
            i = idx % 3
            if i != 0:
                a[idx] = a[idx] * powers[i - 1] % R_MOD

    def coeff_to_extended(self, a: Sequence[int]) -> List[int]:
        """:270-287"""
        a = list(a)
        assert len(a) == 1 << self.k
        self.distribute_powers_zeta(a, True)
        a.extend([0] * (self.extended_len() - len(a)))  # :280
        best_fft(a, self.extended_omega, self.extended_k)
        return a

    def extended_to_coeff(self, a: Sequence[int]) -> List[int]:
        """:328-350"""
        a = list(a)
        assert len(a) == self.extended_len()
        self.ifft(a, self.extended_omega_inv, self.extended_k, self.extended_ifft_divisor)
        self.distribute_powers_zeta(a, False)
        return a[: self.n * self.quotient_poly_degree]  # :346-347

    def divide_by_vanishing_poly(self, a: Sequence[int]) -> List[int]:
        """:354-373"""
        a = list(a)
        assert len(a) == self.extended_len()
        m = len(self.t_evaluations)
        return [v * self.t_evaluations[i % m] % R_MOD for i, v in enumerate(a)]

    def rotate_omega(self, value: int, rotation: int) -> int:
        if rotation < 0:
            return value * pow(self.omega_inv, -rotation, R_MOD) % R_MOD
        return value * pow(self.omega, rotation, R_MOD) % R_MOD

    def l_i_range(self, x: int, xn: int, rotations: Iterable[int]) -> List[int]:
        """poly/domain.rs:497-522 (barycentric form)."""
        rotations = list(rotations)
        denoms = [(x - self.rotate_omega(1, rot)) % R_MOD for rot in rotations]
        common = (xn - 1) * self.barycentric_weight % R_MOD
        return [fr_inv(d) * common % R_MOD * self.rotate_omega(1, rot) % R_MOD
                for d, rot in zip(denoms, rotations)]


def eval_polynomial(poly: Sequence[int], point: int) -> int:
    """arithmetic.rs:707-711 (Horner)."""
    acc = 0
    for c in reversed(poly):
        acc = (acc * point + c) % R_MOD
    return acc


def kate_division(a: Sequence[int], b: int) -> List[int]:
    """arithmetic.rs:752-773: a(X) divided by (X - b), remainder dropped."""
    b = (-b) % R_MOD  # :758
    q = [0] * (len(a) - 1)
    tmp = 0
    for j in range(len(q) - 1, -1, -1):  # :764-770 (q and a walked from the high end)
        lead = (a[j + 1] - tmp) % R_MOD
        q[j] = lead
        tmp = lead * b % R_MOD
    return q


def lagrange_interpolate(points: Sequence[int], evals: Sequence[int]) -> List[int]:
    """arithmetic.rs:848-906 restated as plain Lagrange interpolation."""
    n = len(points)
    coeffs = [0] * n
    for j in range(n):
        num = [1]
        den = 1
        for m in range(n):
            if m == j:
                continue
            new = [0] * (len(num) + 1)
            for i, c in enumerate(num):
                new[i] = (new[i] - c * points[m]) % R_MOD
                new[i + 1] = (new[i + 1] + c) % R_MOD
            num = new
            den = den * (points[j] - points[m]) % R_MOD
        scale = evals[j] * fr_inv(den) % R_MOD
        for i, c in enumerate(num):
            coeffs[i] = (coeffs[i] + c * scale) % R_MOD
    return coeffs


# --------------------------------------------------------------------------
# Params: restatement of poly/commitment.rs
# --------------------------------------------------------------------------
def fq_sqrt(a: int) -> Optional[int]:
    """q = 3 mod 4: a^((q+1)/4) when a is a square"""
    y = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    return y if y * y % Q_MOD == a % Q_MOD else None


def g1_to_bytes(p: Point, sign_bit: int = 7) -> bytes:
    """GroupEncoding::to_bytes of the pinned pairing crate.  [EXT]: restated from the pasta /
    pairing_bn256 convention -- x little-endian canonical, parity of canonical y in the top bit of byte 31,
    identity = 32 zero bytes.  Byte-level parity with the reference's params files is unpinned."""
    if p is None:
        return bytes(32)
    x, y = p
    b = bytearray(x.to_bytes(32, "little"))
    b[31] |= (y & 1) << sign_bit
    return bytes(b)


def g1_from_bytes(b: bytes, sign_bit: int = 7) -> Point:
    """GroupEncoding::from_bytes (same convention); raises ValueError where the reference's
    `Option::from(C::from_bytes(..)).unwrap()` panics (poly/commitment.rs:270)."""
    assert len(b) == 32
    t = bytearray(b)
    sign = (t[31] >> sign_bit) & 1
    t[31] &= ~(1 << sign_bit) & 0xFF
    x = int.from_bytes(bytes(t), "little")
    if x == 0 and sign == 0:
        return None
    if x >= Q_MOD:
        raise ValueError("x is not canonical")
    y = fq_sqrt((x * x * x + CURVE_B) % Q_MOD)
    if y is None:
        raise ValueError("not on the curve")
    if (y & 1) != sign:
        y = Q_MOD - y
    return (x, y)


class Params:
    """poly/commitment.rs:23-29, unsafe_setup :56-124, commit* :129-222, write / read :241-294."""

    def write(self, additional_data: bytes = b"", sign_bit: int = 7) -> bytes:
        """:241-253"""
        out = bytearray(self.k.to_bytes(4, "little"))
        for p in self.g:
            out += g1_to_bytes(p, sign_bit)
        for p in self.g_lagrange:
            out += g1_to_bytes(p, sign_bit)
        out += len(additional_data).to_bytes(4, "little") + additional_data
        return bytes(out)

    @classmethod
    def read(cls, data: bytes, sign_bit: int = 7):
        """:256-294 -> (k, g, g_lagrange, additional_data)"""
        k = int.from_bytes(data[:4], "little")
        n = 1 << k
        pts = [g1_from_bytes(data[4 + 32 * i: 36 + 32 * i], sign_bit) for i in range(2 * n)]
        off = 4 + 64 * n
        ln = int.from_bytes(data[off:off + 4], "little")
        return k, pts[:n], pts[n:], data[off + 4: off + 4 + ln]

    def __init__(self, k: int, s: int):
        assert k <= FR_S  # :60
        self.k = k
        self.n = 1 << k
        n = self.n
        # g[i] = [s^i] G  (:63-83)
        self.g: List[Point] = []
        cur = 1
        for _ in range(n):
            self.g.append(g1_mul(G1_GEN, cur))
            cur = cur * s % R_MOD
        # g_lagrange[i] = [ (s^n - 1)/n * w^i / (s - w^i) ] G  (:85-112)
        root = FR_ROOT_OF_UNITY
        for _ in range(k, FR_S):
            root = root * root % R_MOD
        n_inv = fr_inv(n % R_MOD)
        multiplier = (pow(s, n, R_MOD) - 1) * n_inv % R_MOD
        self.g_lagrange: List[Point] = []
        for i in range(n):
            root_pow = pow(root, i, R_MOD)
            scalar = multiplier * root_pow % R_MOD * fr_inv((s - root_pow) % R_MOD) % R_MOD
            self.g_lagrange.append(g1_mul(G1_GEN, scalar))

    def commit(self, poly: Sequence[int]) -> Point:
        """:129-133"""
        size = len(poly)
        assert len(self.g) >= size
        return best_multiexp_gpu_cond(poly, self.g[:size])

    def commit_lagrange(self, poly: Sequence[int]) -> Point:
        """:138-142"""
        size = len(poly)
        assert len(self.g) >= size
        return best_multiexp_gpu_cond(poly, self.g_lagrange[:size])

    def commit_lagrange_with_bound(self, poly: Sequence[int], _max_bits: int) -> Point:
        """:199-222: drop zero scalars (and their bases), then MSM."""
        scalars, bases = [], []
        for s, b in zip(poly, self.g_lagrange):
            if s % R_MOD != 0:
                scalars.append(s)
                bases.append(b)
        return best_multiexp_gpu_cond(scalars, bases)

    def commit_lagrange_and_ifft(self, poly: Sequence[int], omega_inv: int, ifft_divisor: int):
        """:176-197 (non-cuda): commit_lagrange, then best_fft + scale."""
        c = self.commit_lagrange(poly)
        a = list(poly)
        best_fft(a, omega_inv, self.k)
        a = [v * ifft_divisor % R_MOD for v in a]
        return a, c
