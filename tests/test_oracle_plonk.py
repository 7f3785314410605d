"""Pins oracle/plonk.py (CPU only): the evaluate_h restatement must produce a numerator that the
vanishing polynomial divides with a quotient of degree < n * (degree - 1) for a satisfied circuit --
the identity behind the reference's prove -> verify tests -- and must not when a constraint breaks."""
import pytest

from oracle import bn254 as o
from oracle import plonk as P

import plonk_fixture as fxm


@pytest.fixture(scope="module")
def good():
    fx = fxm.build(k=5, seed=7, domain_j=9)
    return fx, fxm.oracle_h(fx)


def test_satisfied_circuit_is_divisible(good):
    fx, h = good
    assert any(v != 0 for v in h)
    high = fxm.quotient_high_coefficients(fx, h)
    assert len(high) >= 4 * fx["n"] and all(v == 0 for v in high)


@pytest.mark.parametrize("what", ["witness", "copy", "lookup", "shuffle"])
def test_broken_constraint_is_not_divisible(what):
    fx = fxm.build(k=5, seed=7, break_what=what, domain_j=9)
    assert any(v != 0 for v in fxm.quotient_high_coefficients(fx, fxm.oracle_h(fx)))


def test_quotient_identity_at_random_point(good):
    """h(x) * (x^n - 1) == fold of all terms at x, with every polynomial evaluated from its coefficients
    (what the verifier recomputes, plonk/verifier.rs): cross-checks evaluate_h against eval_polynomial."""
    fx, h = good
    d = fx["domain"]
    hq = d.extended_to_coeff(d.divide_by_vanishing_poly(h))
    assert len(hq) == d.n * d.quotient_poly_degree
    assert all(v == 0 for v in hq[4 * d.n - 4:])
    # evaluate the numerator at x through a second, independent route: extended evaluations at a domain point
    # are h itself, so pick an extended-domain row and compare h_quotient(x) * (x^n - 1) with h[row]
    row = 37
    x = d.g_coset * pow(d.extended_omega, row, o.R_MOD) % o.R_MOD
    lhs = o.eval_polynomial(hq, x) * ((pow(x, d.n, o.R_MOD) - 1) % o.R_MOD) % o.R_MOD
    assert lhs == h[row]


def test_evaluator_cse_and_constants():
    cs = P.ConstraintSystem(1, 2, 0, degree=3)
    a, b = P.Advice(0), P.Advice(1)
    cs.gates.append([P.Prod(a, b), P.Prod(a, b), P.Sum(P.Prod(a, b), P.Const(0)), P.Scaled(a, 1), P.Scaled(b, 0)])
    ev = P.Evaluator.new(cs)
    assert ev.constants[:2] == [0, 1]
    # a*b appears once; the three uses share it; Scaled(a, 1) is a itself; Scaled(b, 0) is Constant(0)
    muls = [c for c in ev.calculations if c[0] == "Mul"]
    assert len(muls) == 1
    assert ev.value_parts[0] == ev.value_parts[1] == ev.value_parts[2]
    assert ev.value_parts[4] == ("Constant", 0)


def test_grand_products_close(good):
    fx, _ = good
    cs, n = fx["cs"], fx["n"]
    u = n - (cs.blinding_factors() + 1)
    assert fx["perm_z"][0][0] == 1 and fx["perm_z"][-1][u] == 1          # permutation/prover.rs sanity
    for lk in fx["lookups_lagrange"]:
        assert lk["z"][0][0] == 0 and lk["z"][-1][u] == 0               # logup/prover.rs:364-368
    for z in fx["shuffle_z"]:
        assert z[0] == 1 and z[u] == 1                                   # shuffle/prover.rs:125,152
