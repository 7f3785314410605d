"""BN254 optimal-ate pairing in plain Python big ints -- what Decider::verify needs
(/root/reference/halo2_proofs/src/poly/multiopen.rs:31-57: multi_miller_loop + final_exponentiation + is_identity).

TEST INFRASTRUCTURE ONLY (oracle rules, see oracle/bn254.py).  The pairing itself lives in the reference's
third-party dependency (pairing_bn256 0.1.1 @ lanbones/pairing 30b052f2, absent from /root/reference), so this is a
restatement of the published construction, not of reference source:
  Fq2 = Fq[i]/(i^2 + 1),  xi = 9 + i,  Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)  (so w^6 = xi),
  E : y^2 = x^3 + 3 over Fq,  twist E' : y^2 = x^3 + 3/xi over Fq2,  psi(x, y) = (x w^2, y w^3),
  e(P, Q) = f_{6t+2,Q}(P) * l_{[6t+2]Q, pi(Q)}(P) * l_{[6t+2]Q + pi(Q), -pi^2(Q)}(P), raised to (q^12 - 1)/r,
  with t = 4965661367192848881 (6t + 2 = 29793968203157093288).
Pinned by bilinearity and non-degeneracy (tests/test_oracle_prover.py): e([a]P, Q) == e(P, [a]Q) != 1, r * G2 == O.
A product of pairings being 1 does not depend on the choice of generator or on the normalisation of the Miller
function, which is all Decider::verify uses.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .bn254 import Q_MOD as Q, R_MOD as R

Fq2 = Tuple[int, int]
G2Point = Optional[Tuple[Fq2, Fq2]]

ATE_LOOP_COUNT = 29793968203157093288
XI: Fq2 = (9, 1)


# ------------------------------------------------------------------ Fq2
def f2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a: Fq2) -> Fq2:
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_scalar(a: Fq2, k: int) -> Fq2:
    return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a: Fq2) -> Fq2:
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


def f2_conj(a: Fq2) -> Fq2:
    return (a[0], (-a[1]) % Q)


def f2_pow(a: Fq2, e: int) -> Fq2:
    out: Fq2 = (1, 0)
    while e:
        if e & 1:
            out = f2_mul(out, a)
        a = f2_mul(a, a)
        e >>= 1
    return out


B2: Fq2 = f2_scalar(f2_inv(XI), 3)                      # twist coefficient 3 / xi
GAMMA2 = f2_pow(XI, (Q - 1) // 3)                       # w^(2(q-1))
GAMMA3 = f2_pow(XI, (Q - 1) // 2)                       # w^(3(q-1))

G2_GEN: G2Point = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ------------------------------------------------------------------ G2 (affine over Fq2)
def g2_is_on_curve(p: G2Point) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


def g2_neg(p: G2Point) -> G2Point:
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_add(p: G2Point, q: G2Point) -> G2Point:
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if y1 != y2 or y1 == (0, 0):
            return None
        m = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(m, m), x1), x2)
    return (x3, f2_sub(f2_mul(m, f2_sub(x1, x3)), y1))


def g2_mul(p: G2Point, k: int) -> G2Point:
    k %= R
    acc: G2Point = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_frobenius(p: G2Point) -> G2Point:
    """pi on the twist: psi^-1 o Frobenius o psi"""
    return (f2_mul(f2_conj(p[0]), GAMMA2), f2_mul(f2_conj(p[1]), GAMMA3))


# ------------------------------------------------------------------ Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)
F12 = List[int]
F12_ONE: F12 = [1] + [0] * 11


def f12_mul(a: F12, b: F12) -> F12:
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                if y:
                    t[i + j] += x * y
    for i in range(22, 11, -1):
        c = t[i]
        if c:
            t[i - 6] += 18 * c
            t[i - 12] -= 82 * c
    return [v % Q for v in t[:12]]


def f12_pow(a: F12, e: int) -> F12:
    out = list(F12_ONE)
    while e:
        if e & 1:
            out = f12_mul(out, a)
        a = f12_mul(a, a)
        e >>= 1
    return out


def _embed(dst: F12, c: Fq2, j: int) -> None:
    """dst += c * w^j for c = a + b i = (a - 9 b) + b w^6, j < 6"""
    dst[j] = (dst[j] + c[0] - 9 * c[1]) % Q
    dst[j + 6] = (dst[j + 6] + c[1]) % Q


def _line(r: G2Point, s: G2Point, p: Tuple[int, int]) -> F12:
    """the line through psi(r), psi(s) (tangent when r == s) evaluated at p in E(Fq):
    -yp + (m xp) w + (y1 - m x1) w^3 with m the slope on the twist; vertical: xp - x1 w^2"""
    (x1, y1), (x2, y2) = r, s
    xp, yp = p
    out = [0] * 12
    if x1 == x2 and y1 != y2:
        out[0] = xp % Q
        _embed(out, f2_neg(x1), 2)
        return out
    if x1 == x2:
        m = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        m = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    out[0] = (-yp) % Q
    _embed(out, f2_scalar(m, xp), 1)
    _embed(out, f2_sub(y1, f2_mul(m, x1)), 3)
    return out


def miller_loop(q: G2Point, p: Optional[Tuple[int, int]]) -> F12:
    if p is None or q is None:
        return list(F12_ONE)
    f = list(F12_ONE)
    r = q
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        f = f12_mul(f12_mul(f, f), _line(r, r, p))
        r = g2_add(r, r)
        if (ATE_LOOP_COUNT >> i) & 1:
            f = f12_mul(f, _line(r, q, p))
            r = g2_add(r, q)
    q1 = g2_frobenius(q)
    nq2 = g2_neg(g2_frobenius(q1))
    f = f12_mul(f, _line(r, q1, p))
    r = g2_add(r, q1)
    f = f12_mul(f, _line(r, nq2, p))
    return f


FINAL_EXP = (Q ** 12 - 1) // R


def final_exponentiation(f: F12) -> F12:
    return f12_pow(f, FINAL_EXP)


def pairing(p, q: G2Point) -> F12:
    return final_exponentiation(miller_loop(q, p))


def pairing_check(pairs: Sequence[Tuple[Optional[Tuple[int, int]], G2Point]]) -> bool:
    """prod e(P_i, Q_i) == 1 (multi_miller_loop, one final exponentiation)"""
    f = list(F12_ONE)
    for p, q in pairs:
        f = f12_mul(f, miller_loop(q, p))
    return final_exponentiation(f) == F12_ONE
