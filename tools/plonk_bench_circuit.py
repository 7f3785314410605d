"""The circuit of the reference's benches/plonk.rs:58-251 (BASELINE config 4), laid out directly: three advice
columns a, b, c with equality enabled, four fixed columns sm, sa, sb, sc, one gate
a*sa + b*sb + a*b*sm - c*sc of degree 3 under set_minimum_degree(5), and 2^(k-1) - 3 iterations of

    row 2i   raw_multiply: (a, b, c) = (x, x, x^2),      sc = sm = 1
    row 2i+1 raw_add:      (a, b, c) = (x, x^2, x + x^2), sa = sb = sc = 1
    copy a[2i] = a[2i+1], copy b[2i+1] = c[2i]

which fill the 2^k - 6 usable rows exactly.  The layouter, selector handling and witness synthesis of the reference
are front-end work outside this repository's scope, so the columns are produced here with numpy.
Returns plain data; no oracle and no engine involved."""
from __future__ import annotations

import numpy as np

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_MONT = (1 << 256) % R


def _mont(v: int) -> np.ndarray:
    v = v % R * _MONT % R
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def expressions():
    A = lambda c: ("Advice", c, 0)       # noqa: E731
    F = lambda c: ("Fixed", c, 0)        # noqa: E731
    a, b, c = A(0), A(1), A(2)
    sm, sa, sb, sc = F(0), F(1), F(2), F(3)
    gate = ("Sum", ("Sum", ("Sum", ("Product", a, sa), ("Product", b, sb)), ("Product", ("Product", a, b), sm)),
            ("Negated", ("Product", c, sc)))
    return gate


def constraint_system_args() -> dict:
    """keyword arguments for halo2_gpu_specific_b200.plonk.ConstraintSystem (query lists in the order of the
    reference's configure(): enable_equality(a, b, c), then a, b, c, sa, sb, sc, sm inside create_gate)"""
    return dict(num_fixed=4, num_advice=3, num_instance=0, degree=5, blinding_factors=5, gates=[[expressions()]],
                permutation_columns=[("Advice", 0), ("Advice", 1), ("Advice", 2)],
                advice_queries=[(0, 0), (1, 0), (2, 0)], fixed_queries=[(1, 0), (2, 0), (3, 0), (0, 0)],
                instance_queries=[])


def build(k: int, x: int = 0x1234567):
    """-> (fixed (4, n, 4), advice (3, n, 4), mapping (3, n, 2)), Montgomery limbs"""
    n = 1 << k
    it = (1 << (k - 1)) - 3
    rows = 2 * it
    assert rows == n - 6
    one, vx, vx2, vs = _mont(1), _mont(x), _mont(x * x), _mont(x + x * x)
    fixed = np.zeros((4, n, 4), dtype=np.uint64)
    advice = np.zeros((3, n, 4), dtype=np.uint64)
    mul_rows = np.arange(0, rows, 2)
    add_rows = mul_rows + 1
    fixed[0, mul_rows] = one                     # sm
    fixed[3, mul_rows] = one                     # sc
    fixed[1, add_rows] = one                     # sa
    fixed[2, add_rows] = one                     # sb
    fixed[3, add_rows] = one                     # sc
    advice[0, :rows] = vx
    advice[1, mul_rows] = vx
    advice[1, add_rows] = vx2
    advice[2, mul_rows] = vx2
    advice[2, add_rows] = vs
    mapping = np.empty((3, n, 2), dtype=np.int64)
    mapping[..., 0] = np.arange(3)[:, None]
    mapping[..., 1] = np.arange(n)[None, :]
    # a[2i] <-> a[2i+1]
    mapping[0, mul_rows, 1] = add_rows
    mapping[0, add_rows, 1] = mul_rows
    # b[2i+1] <-> c[2i]
    mapping[1, add_rows, 0], mapping[1, add_rows, 1] = 2, mul_rows
    mapping[2, mul_rows, 0], mapping[2, mul_rows, 1] = 1, add_rows
    return fixed, advice, mapping
