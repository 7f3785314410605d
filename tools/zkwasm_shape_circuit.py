"""A satisfiable synthetic circuit with the column / argument counts of BASELINE config 5 ("zkWasm-scale
k = 22 circuit with many advice columns, lookups and permutations"): A = 64 advice, F = 32 fixed, I = 1 instance,
L = 8 logup lookups with 12 input sets in total, H = 4 shuffle groups, 24 permutation columns in 8 sets, degree 5 --
the shape tools/proof_replay.py and tools/resident_replay.py replay as a schedule, here with a real witness so that
create_proof produces a proof a verifier accepts.  zkWasm's own gates are not available (the circuit lives in another
repository); the gate set is a stand-in whose size is a parameter (`extra_gates`).

  a[3j], a[3j+1], a[3j+2] = a[3j] * a[3j+1]     j < 15, 16-bit inputs                gate  q_j * (a*b - c)
  a45 16-bit, a46 = running sum of a45                                             gate  q_sum * (a46(wX) - a46 - a45)
  a47 = public inputs in its first rows                                             gate  q_inst * (a47 - instance)
  a48 .. a59 = table[random row]             table t = f16 (distinct values)         8 lookups, 12 single-input sets
  a60 .. a63 = permutations of a48 .. a51                                           4 shuffle groups
  extra gates  q_j * (a*b - c) * (a_x + f_y)    degree 4, satisfied wherever q_j * (a*b - c) is; they sit in the gate of
               their triple, so the shared factor is short-lived in the quotient program (4 live slots instead of 17
               when all extras come last: the y-fold visits the polynomials in gate order)
  copies       a[6i][r] = a[6i + 3][r] for r = 0 mod 4, i < 4                        permutation over a0 .. a23

Everything is generated with numpy on small integers; `to_mont` (canonical (m, 4) limbs -> Montgomery) is supplied by
the caller (the engine's field kernel on the GPU box, the C oracle in CPU tests)."""
from __future__ import annotations

import numpy as np

A, F, I = 64, 32, 1
LOOKUP_SETS = (2, 2, 2, 2, 1, 1, 1, 1)
SHUFFLES = 4
PERM_COLS = 24
TRIPLES = 15
BF = 5
N_PUBLIC = 4


def _adv(c, r=0): return ("Advice", c, r)          # noqa: E704
def _fix(c): return ("Fixed", c, 0)                # noqa: E704
def _mul(a, b): return ("Product", a, b)           # noqa: E704
def _sub(a, b): return ("Sum", a, ("Negated", b))  # noqa: E704
def _add(a, b): return ("Sum", a, b)               # noqa: E704


def constraint_system_args(extra_gates: int = 64, extras_at_end: bool = False) -> dict:
    """extras_at_end: the extra gates follow ALL product gates instead of sitting next to the gate whose factor they
    share -- the same constraints in another gate order, which keeps every shared factor q_j * (a*b - c) alive across
    the whole y-fold (17 live values in the lowered evaluate_h program instead of 5: the case the quotient kernel's
    global slot class is for)"""
    gates = []
    product = lambda j: _sub(_mul(_adv(3 * j), _adv(3 * j + 1)), _adv(3 * j + 2))      # noqa: E731
    late = []
    for j in range(TRIPLES):                          # one gate per triple: its product constraint and its extras, so
        polys = [_mul(_fix(j), product(j))]           # that the shared sub-expression q_j * (a*b - c) is short-lived
        for e in range(j, extra_gates, TRIPLES):
            x, y = (7 * e + 3) % A, 18 + (e % (F - 18))
            (late if extras_at_end else polys).append(_mul(_mul(_fix(j), product(j)), _add(_adv(x), _fix(y))))
        gates.append(polys)
    if late:
        gates.append(late)
    gates.append([_mul(_fix(15), _sub(_sub(_adv(46, 1), _adv(46)), _adv(45)))])
    gates.append([_mul(_fix(17), _sub(_adv(47), ("Instance", 0, 0)))])
    lookups, col = [], 48
    for sets in LOOKUP_SETS:
        lookups.append({"table_expressions": [_fix(16)],
                        "input_expressions_sets": [[[_adv(col + s)]] for s in range(sets)]})
        col += sets
    shuffles = [[{"input_expressions": [_adv(48 + g)], "shuffle_expressions": [_adv(60 + g)]}] for g in range(SHUFFLES)]
    return dict(num_fixed=F, num_advice=A, num_instance=I, degree=5, blinding_factors=BF, gates=gates, lookups=lookups,
                shuffles=shuffles, permutation_columns=[("Advice", c) for c in range(PERM_COLS)])


def build(k: int, to_mont, seed: int = 1):
    """-> (fixed (F, n, 4), advice (A, n, 4), public inputs [ints], mapping (PERM_COLS, n, 2))"""
    n = 1 << k
    usable = n - (BF + 1)
    assert usable > 16
    rng = np.random.Generator(np.random.PCG64(seed))
    u16 = lambda: rng.integers(0, 1 << 16, size=n, dtype=np.uint64)                   # noqa: E731
    rows = np.arange(n)
    adv = np.zeros((A, n), dtype=np.uint64)
    fix = np.zeros((F, n), dtype=np.uint64)
    on = (rows < usable).astype(np.uint64)

    for j in range(TRIPLES):
        adv[3 * j], adv[3 * j + 1] = u16(), u16()
    copy_rows = rows[(rows % 4 == 0) & (rows < usable)]
    for i in range(4):                                              # copies between input cells, before the products
        adv[6 * i + 3, copy_rows] = adv[6 * i, copy_rows]
    for j in range(TRIPLES):
        adv[3 * j + 2] = adv[3 * j] * adv[3 * j + 1]
        fix[j] = on
    adv[45] = u16()
    adv[46, 1:] = np.cumsum(adv[45, :-1], dtype=np.uint64)
    fix[15] = (rows < usable - 1).astype(np.uint64)
    adv[47] = u16()
    public = [int(v) for v in adv[47, :N_PUBLIC]]
    fix[17] = (rows < N_PUBLIC).astype(np.uint64)
    table = rows.astype(np.uint64) * np.uint64(7) + np.uint64(3)
    fix[16] = table * on
    for j in range(12):
        adv[48 + j] = table[rng.integers(0, usable, size=n)]
    for g in range(SHUFFLES):
        perm = rng.permutation(usable)
        adv[60 + g, :usable] = adv[48 + g, perm]
    for y in range(18, F):
        fix[y] = u16()

    def limbs(cols: np.ndarray) -> np.ndarray:
        out = np.empty(cols.shape + (4,), dtype=np.uint64)
        one = np.zeros((n, 4), dtype=np.uint64)
        for c in range(cols.shape[0]):                              # column by column: bounded staging memory
            one[:, 0] = cols[c]
            out[c] = to_mont(one).reshape(n, 4)
        return out

    mapping = np.empty((PERM_COLS, n, 2), dtype=np.int64)
    mapping[..., 0] = np.arange(PERM_COLS)[:, None]
    mapping[..., 1] = rows[None, :]
    for i in range(4):                                              # 2-cycles (6i, r) <-> (6i + 3, r)
        mapping[6 * i, copy_rows, 0] = 6 * i + 3
        mapping[6 * i + 3, copy_rows, 0] = 6 * i
    return limbs(fix), limbs(adv), public, mapping
