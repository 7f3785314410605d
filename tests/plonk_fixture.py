"""A small satisfied PLONK instance built with the oracle (oracle/plonk.py): custom gates with
positive / negative rotations, constants and scaling, a permutation over advice / fixed / instance
columns split into two column chunks, two logup lookups (two input sets, theta-compressed pairs),
and a shuffle group of two arguments.  Shared by the oracle tests and the GPU parity tests."""
from __future__ import annotations

import random

from oracle import bn254 as o
from oracle import plonk as P

R = o.R_MOD


def build(k: int = 5, seed: int = 1, break_what: str | None = None, domain_j: int | None = None):
    rng = random.Random(seed)
    n = 1 << k
    cs = P.ConstraintSystem(num_fixed=6, num_advice=9, num_instance=1, degree=5, blinding_factors=5)
    bf = cs.blinding_factors()
    usable = n - bf - 1
    assert usable >= 26
    A, F, I = P.Advice, P.Fixed, P.Instance
    q_mul, q_add, q_rot, q_neg = F(0), F(1), F(2), F(3)
    # gates
    cs.gates.append([P.Prod(q_mul, P.Sub(P.Prod(A(0), A(1)), A(2)))])
    cs.gates.append([P.Prod(q_add, P.Sub(P.Sum(A(0), A(1)), A(2))),
                     P.Prod(q_add, P.Prod(A(0), P.Sub(A(0), A(0))))])          # second polynomial of the gate: identically 0
    cs.gates.append([P.Prod(q_rot, P.Sub(A(0, 1), P.Sum(A(2), I(0))))])
    cs.gates.append([P.Prod(q_neg, P.Sub(A(2), P.Sum(P.Scaled(A(2, -1), 7), P.Const(5))))])
    # permutation
    cs.permutation_columns = [("Advice", 0), ("Advice", 1), ("Advice", 2), ("Fixed", 5), ("Instance", 0)]
    # lookups: table columns fixed 4 (t0) and, for the pair lookup, (t0, t1 = fixed 5)
    cs.lookups.append({"table_expressions": [F(4)], "input_expressions_sets": [[[A(5)], [A(6)]], [[A(7)]]]})
    cs.lookups.append({"table_expressions": [F(4), F(5)], "input_expressions_sets": [[[A(5), A(6)]]]})
    # shuffle group with two arguments
    cs.shuffles.append([{"input_expressions": [A(3)], "shuffle_expressions": [A(4)]},
                        {"input_expressions": [A(5)], "shuffle_expressions": [A(8)]}])

    # ---- fixed columns
    fixed = [[0] * n for _ in range(cs.num_fixed)]
    for r in range(0, 10):
        fixed[0][r] = 1
    for r in range(10, 18):
        fixed[1][r] = 1
    for r in range(18, 22):
        fixed[2][r] = 1
    for r in range(23, 26):
        fixed[3][r] = 1
    t0 = [(i * i + 3) % R for i in range(n)]
    fixed[4] = list(t0)
    fixed[5] = [t0[(i + 5) % usable] if i < usable else rng.randrange(R) for i in range(n)]
    instance = [[rng.randrange(R) if r < 4 else 0 for r in range(n)]]

    # ---- advice
    adv = [[rng.randrange(R) for _ in range(n)] for _ in range(cs.num_advice)]
    copies = [(("Advice", 0, 3), ("Advice", 2, 0)), (("Advice", 1, 4), ("Advice", 2, 1)),
              (("Advice", 0, 12), ("Advice", 2, 2)), (("Advice", 1, 12), ("Fixed", 5, 0)),
              (("Advice", 0, 13), ("Instance", 0, 0))]

    def cell(kind, idx, row):
        return {"Advice": adv, "Fixed": fixed, "Instance": instance}[kind][idx][row]

    for r in range(usable):
        for (dk, di, dr), (sk, si, sr) in copies:
            if dr == r:
                assert dk == "Advice"
                adv[di][r] = cell(sk, si, sr)
        if r >= 1 and fixed[2][r - 1]:
            adv[0][r] = (adv[2][r - 1] + instance[0][r - 1]) % R
        if fixed[0][r]:
            adv[2][r] = adv[0][r] * adv[1][r] % R
        elif fixed[1][r]:
            adv[2][r] = (adv[0][r] + adv[1][r]) % R
        elif fixed[3][r]:
            adv[2][r] = (7 * adv[2][r - 1] + 5) % R
    # lookup inputs: (a5, a6) is a row of (t0, t1); a7 in t0
    for r in range(usable):
        j = rng.randrange(usable)
        adv[5][r], adv[6][r] = fixed[4][j], fixed[5][j]
        adv[7][r] = fixed[4][rng.randrange(usable)]
    # shuffles: a4 = permutation of a3, a8 = permutation of a5 (usable rows)
    perm = list(range(usable))
    rng.shuffle(perm)
    for r in range(usable):
        adv[4][r] = adv[3][perm[r]]
    rng.shuffle(perm)
    for r in range(usable):
        adv[8][r] = adv[5][perm[r]]

    mapping = P.identity_mapping(len(cs.permutation_columns), n)
    colpos = {c: i for i, c in enumerate(cs.permutation_columns)}
    for (dk, di, dr), (sk, si, sr) in copies:
        P.mapping_copy(mapping, (colpos[(dk, di)], dr), (colpos[(sk, si)], sr))

    if break_what == "witness":
        adv[2][4] = (adv[2][4] + 1) % R
    elif break_what == "copy":
        adv[0][3] = (adv[0][3] + 1) % R
        adv[2][3] = adv[0][3] * adv[1][3] % R      # the gate still holds, the copy does not
    elif break_what == "shuffle":
        adv[4][2] = (adv[4][2] + 1) % R

    # domain_j > cs.degree() gives a larger extended domain than the quotient needs, which makes
    # "the quotient has degree < n * (degree - 1)" a checkable statement (see quotient_high_coefficients)
    domain = o.EvaluationDomain(domain_j or cs.degree(), k)
    theta, beta, gamma, y = (rng.randrange(R) for _ in range(4))

    sigmas = P.permutation_sigmas(cs, domain, mapping)
    perm_z = P.permutation_commit(cs, domain, sigmas, adv, fixed, instance, beta, gamma, rng)
    lookups_lagrange = []
    for lk in cs.lookups:
        input_sets, table, m = P.logup_compress(cs, domain, lk, theta, adv, fixed, instance, rng)
        if break_what == "lookup" and lk is cs.lookups[0]:
            m[0] = (m[0] + 1) % R
        zs = [P.blind_to_n(z, n, rng) for z in P.logup_commit_z(cs, domain, input_sets, table, m, beta)]
        lookups_lagrange.append({"z": zs, "m": m, "input_sets": input_sets, "table": table})
    shuffle_z = [P.blind_to_n(P.shuffle_commit_product(cs, domain, g, theta, beta, adv, fixed, instance), n, rng)
                 for g in cs.shuffles]

    to_coeff = domain.lagrange_to_coeff
    coeff = {
        "fixed": [to_coeff(c) for c in fixed],
        "advice": [to_coeff(c) for c in adv],
        "instance": [to_coeff(c) for c in instance],
        "sigma": [to_coeff(c) for c in sigmas],
        "perm_z": [to_coeff(c) for c in perm_z],
        "lookup_z": [[to_coeff(z) for z in lk["z"]] for lk in lookups_lagrange],
        "lookup_m": [to_coeff(lk["m"]) for lk in lookups_lagrange],
        "shuffle_z": [to_coeff(z) for z in shuffle_z],
    }
    return {
        "k": k, "n": n, "cs": cs, "domain": domain, "ev": P.Evaluator.new(cs),
        "fixed": fixed, "advice": adv, "instance": instance, "sigmas": sigmas, "mapping": mapping,
        "perm_z": perm_z, "lookups_lagrange": lookups_lagrange, "shuffle_z": shuffle_z,
        "coeff": coeff, "theta": theta, "beta": beta, "gamma": gamma, "y": y,
    }


def cosets(fx):
    """extended-domain evaluations of every polynomial evaluate_h reads (oracle transforms)"""
    d = fx["domain"]
    ext = d.coeff_to_extended
    c = fx["coeff"]
    l0, l_last, l_active = P.lagrange_basis_cosets(fx["cs"], d)
    return {
        "fixed": [ext(p) for p in c["fixed"]],
        "advice": [ext(p) for p in c["advice"]],
        "instance": [ext(p) for p in c["instance"]],
        "sigma": [ext(p) for p in c["sigma"]],
        "perm_z": [ext(p) for p in c["perm_z"]],
        "lookups": [{"z_cosets": [ext(z) for z in zs], "m_coset": ext(m)}
                    for zs, m in zip(c["lookup_z"], c["lookup_m"])],
        "shuffles": [ext(p) for p in c["shuffle_z"]],
        "l0": l0, "l_last": l_last, "l_active_row": l_active,
    }


def oracle_h(fx, cz=None):
    cz = cz or cosets(fx)
    return P.evaluate_h(fx["ev"], fx["cs"], fx["domain"], cz["fixed"], cz["advice"], cz["instance"], cz["l0"],
                        cz["l_last"], cz["l_active_row"], cz["sigma"], fx["y"], fx["beta"], fx["gamma"], fx["theta"],
                        cz["lookups"], cz["shuffles"], cz["perm_z"])


def quotient_high_coefficients(fx, h):
    """h / (X^n - 1) over the extended coset, back to coefficients WITHOUT truncation.  Every term of
    the numerator has degree <= degree * (n - 1), so for a satisfied circuit the quotient's coefficients
    from index (degree - 1) * n - (degree - 1) on are exactly zero; the list is only non-empty when the
    fixture was built with domain_j > degree (extended domain larger than the quotient needs)."""
    d = fx["domain"]
    deg = fx["cs"].degree()
    q = d.divide_by_vanishing_poly(h)
    a = list(q)
    d.ifft(a, d.extended_omega_inv, d.extended_k, d.extended_ifft_divisor)
    d.distribute_powers_zeta(a, False)
    return a[(deg - 1) * d.n - (deg - 1):]
