"""Randomised circuits through the whole prover (CPU): random gate expressions (constants, negation, scaling, sums,
products, rotations) made satisfiable by construction, a random permutation, a lookup and a shuffle.  For every seed
the oracle's proof must verify and the prover mirror's host logic (over the oracle-backed engine) must write the same
bytes -- this walks the corners of Evaluator::add_expression (constant folding, the `0 - b` quirk, operand ordering,
sub-expression sharing) and of the query bookkeeping that the hand-written fixture does not reach."""
import random

import numpy as np
import pytest

from oracle import bn254 as o
from oracle import plonk as P
from oracle import prover as PR
from oracle_engine import OracleEngine

from halo2_gpu_specific_b200 import plonk as HP

R = o.R_MOD
enc = o.fr_encode
S_TOXIC = 0x2B200B200B200B200B200B200B200B2001
K = 5
N = 1 << K
BF = 5
USABLE = N - BF - 1
N_IN = 4            # free input columns a0 .. a3
N_GATES = 3         # target columns a4 .. a6
N_FIXED_CONST = 2   # f0, f1: random fixed columns usable inside expressions


class HostParams:
    k, n = K, N


def random_expression(rng, depth, budget):
    """a random Expression of degree <= budget over the input columns and the constant fixed columns"""
    if depth == 0 or budget == 0 or rng.random() < 0.2:
        if budget == 0 or rng.random() < 0.25:
            return P.Const(rng.choice([0, 1, 2, R - 1, rng.randrange(R)]))
        if rng.random() < 0.3:
            return P.Fixed(rng.randrange(N_FIXED_CONST), rng.choice([0, 0, 1, -1]))
        return P.Advice(rng.randrange(N_IN), rng.choice([0, 0, 0, 1, -1, 2, -2]))
    kind = rng.choice(["neg", "scaled", "sum", "sub", "prod", "prod"])
    if kind == "neg":
        return P.Neg(random_expression(rng, depth - 1, budget))
    if kind == "scaled":
        return P.Scaled(random_expression(rng, depth - 1, budget), rng.choice([0, 1, 3, rng.randrange(R)]))
    if kind == "sum":
        return P.Sum(random_expression(rng, depth - 1, budget), random_expression(rng, depth - 1, budget))
    if kind == "sub":
        return P.Sub(random_expression(rng, depth - 1, budget), random_expression(rng, depth - 1, budget))
    left = rng.randint(0, budget)
    return P.Prod(random_expression(rng, depth - 1, left), random_expression(rng, depth - 1, budget - left))


def folds_to_zero(e) -> bool:
    """does Evaluator::add_expression reduce e to Constant(0)?  (evaluation.rs:671-776: a zero constant, Scaled by 0,
    a zero factor, 0 + 0, ...: asked of the restated evaluator itself)"""
    ev = P.Evaluator()
    ev.add_constant(0)
    ev.add_constant(1)
    return ev.add_expression(e) == ("Constant", 0)


def has_zero_minus(e) -> bool:
    """`0 - b`: the reference's add_expression returns b instead of -b for it (evaluation.rs:727-728), so the prover's
    numerator disagrees with the verifier's expression and such a circuit cannot be proven -- see the dedicated test"""
    if e[0] == "Sum" and e[2][0] == "Negated" and folds_to_zero(e[1]):
        return True
    return any(has_zero_minus(c) for c in e[1:] if isinstance(c, tuple) and c and isinstance(c[0], str))


def rotations_of(e, acc):
    if e[0] in ("Fixed", "Advice", "Instance"):
        acc.add(e[2])
    elif e[0] in ("Negated", "Scaled"):
        rotations_of(e[1], acc)
    elif e[0] in ("Sum", "Product"):
        rotations_of(e[1], acc)
        rotations_of(e[2], acc)
    return acc


def build(seed):
    rng = random.Random(seed)
    # lookup shape: 1-2 input sets of 1-2 inputs, each input a tuple of 1-2 expressions (theta-compressed)
    lk_width = rng.randint(1, 2)
    lk_sets = [rng.randint(1, 2) for _ in range(rng.randint(1, 2))]
    n_lk_cols = sum(lk_sets) * lk_width
    n_adv = N_IN + N_GATES + n_lk_cols + 4         # + lookup inputs, shuffle input, shuffle column, 2 copy columns
    n_fix = N_FIXED_CONST + N_GATES + 2            # + one selector per gate + two table columns
    cs = P.ConstraintSystem(n_fix, n_adv, 1, degree=5, blinding_factors=BF)
    fixed = [[rng.randrange(R) for _ in range(N)] for _ in range(n_fix)]
    advice = [[rng.randrange(R) for _ in range(N)] for _ in range(n_adv)]
    instance = [[rng.randrange(R) if r < 3 else 0 for r in range(N)]]
    for g in range(N_GATES):
        e = random_expression(rng, 3, 3)
        while has_zero_minus(e) or folds_to_zero(e):     # the gate is q * (e - target): e = 0 would be `0 - b` itself
            e = random_expression(rng, 3, 3)
        if g == 0:
            e = P.Sum(e, P.Instance(0, 0))         # keep the instance column in play
        rots = rotations_of(e, {0})
        sel = N_FIXED_CONST + g
        tgt = N_IN + g
        for r in range(N):
            ok = all(0 <= r + t < USABLE for t in rots)
            fixed[sel][r] = 1 if ok else 0
            if ok:
                advice[tgt][r] = P.eval_expr(e, r, N, 1, fixed, advice, instance)
        cs.gates.append([P.Prod(P.Fixed(sel), P.Sub(e, P.Advice(tgt)))])
    table = N_FIXED_CONST + N_GATES
    lk_first = N_IN + N_GATES
    sh_in, sh_out = lk_first + n_lk_cols, lk_first + n_lk_cols + 1
    repeat = rng.choice([1, 1, 2, 3])               # repeated table rows: the multiplicity goes to the searched row
    fixed[table] = [(7 * (i // repeat) + 1) % R for i in range(N)]
    fixed[table + 1] = [(11 * (i // repeat) + 5) % R for i in range(N)]
    table_exprs = [P.Fixed(table + w) for w in range(lk_width)]
    sets, col = [], lk_first
    for n_inputs in lk_sets:
        inputs = []
        for _ in range(n_inputs):
            for r in range(USABLE):
                j = rng.randrange(USABLE)
                for w in range(lk_width):
                    advice[col + w][r] = fixed[table + w][j]
            inputs.append([P.Advice(col + w) for w in range(lk_width)])
            col += lk_width
        sets.append(inputs)
    perm = list(range(USABLE))
    rng.shuffle(perm)
    for r in range(USABLE):
        advice[sh_out][r] = advice[sh_in][perm[r]]
    cs.lookups.append({"table_expressions": table_exprs, "input_expressions_sets": sets})
    cs.shuffles.append([{"input_expressions": [P.Advice(sh_in)], "shuffle_expressions": [P.Advice(sh_out)]}])
    # permutation over two advice columns no gate reads, a fixed column and the instance column: random cycles of two
    # to four cells, some anchored at a fixed or an instance cell (whose value the advice cells must then take)
    p0, p1 = sh_out + 1, sh_out + 2
    cs.permutation_columns = [("Advice", p0), ("Advice", p1), ("Fixed", 0), ("Instance", 0)]
    mapping = P.identity_mapping(4, N)
    free = [(c, r) for c in (0, 1) for r in range(USABLE)]
    rng.shuffle(free)
    for _ in range(rng.randint(2, 6)):
        kind = rng.choice(["advice", "advice", "fixed", "instance"])
        if kind == "fixed":
            anchor, value = (2, rng.randrange(USABLE)), None
            value = fixed[0][anchor[1]]
        elif kind == "instance":
            anchor = (3, rng.randrange(3))
            value = instance[0][anchor[1]]
        else:
            anchor = free.pop()
            value = advice[(p0, p1)[anchor[0]]][anchor[1]]
        for _cell in range(rng.randint(1, 3)):
            c, r = free.pop()
            advice[(p0, p1)[c]][r] = value
            P.mapping_copy(mapping, anchor, (c, r))
    return cs, fixed, advice, instance, mapping


@pytest.mark.parametrize("seed", range(24))
def test_random_circuit(seed):
    cs, fixed, advice, instance, mapping = build(seed)
    oparams = PR.Params(K, S_TOXIC)
    opk = PR.keygen(oparams, cs, fixed, mapping)
    inst = [instance[0][:3]]
    use_gwc = seed % 2 == 0
    proof = PR.create_proof(oparams, opk, advice, inst, HP.SeededRng(seed), use_gwc=use_gwc)
    assert PR.verify_proof(oparams, opk.vk, inst, proof, use_gwc=use_gwc), "oracle verifier rejects a satisfied circuit"
    hcs = HP.ConstraintSystem.like(cs)
    parts = HP.evaluator_parts(hcs)
    assert parts["calculations"] == opk.ev.calculations and parts["constants"] == opk.ev.constants
    assert parts["value_parts"] == opk.ev.value_parts and parts["rotations"] == opk.ev.rotations
    eng = OracleEngine(oparams, opk.vk.domain, cs)
    pk = HP.keygen(HostParams, hcs, np.stack([enc(c) for c in fixed]), np.array(mapping, dtype=np.int64), engine=eng,
                   transcript_repr=opk.vk.transcript_repr)
    got = HP.create_proof(HostParams, pk, np.ascontiguousarray(np.stack([enc(c) for c in advice])), inst,
                          HP.SeededRng(seed), engine=eng, use_gwc=use_gwc)
    assert got == proof
    # break one copied cell (a cell whose permutation image is not itself): the proof must be rejected
    moved = [(c, r) for c in (0, 1) for r in range(USABLE) if tuple(mapping[c][r]) != (c, r)]
    if moved:
        c, r = moved[0]
        col = cs.permutation_columns[c][1]
        bad = [list(x) for x in advice]
        bad[col][r] = (bad[col][r] + 1) % R
        assert not PR.verify_proof(oparams, opk.vk, inst, PR.create_proof(oparams, opk, bad, inst, HP.SeededRng(seed),
                                                                        use_gwc=use_gwc), use_gwc=use_gwc)
    # break one gate target: the proof must be rejected
    bad = [list(c) for c in advice]
    rows_on = [r for r in range(N) if fixed[N_FIXED_CONST][r]]
    if rows_on:
        bad[N_IN][rows_on[0]] = (bad[N_IN][rows_on[0]] + 1) % R
        assert not PR.verify_proof(oparams, opk.vk, inst, PR.create_proof(oparams, opk, bad, inst, HP.SeededRng(seed),
                                                                        use_gwc=use_gwc), use_gwc=use_gwc)


def test_zero_minus_b_cannot_be_proven():
    """The quirk restated from the reference (evaluation.rs:727-728: `0 - b` is lowered to `b`): for a witness that
    satisfies the gate as written, the prover's numerator is not divisible by the vanishing polynomial, so the
    verifier -- which evaluates the expression as written -- rejects.  Both the oracle and the prover mirror lower the
    expression the reference's way; the point of this test is that the restatement keeps the quirk instead of fixing it."""
    cs = P.ConstraintSystem(1, 2, 0, degree=3, blinding_factors=BF)
    cs.gates.append([P.Prod(P.Fixed(0), P.Sub(P.Sub(P.Const(0), P.Advice(0)), P.Advice(1)))])     # q * ((0 - a) - t)
    rng = random.Random(1)
    a = [rng.randrange(R) for _ in range(N)]
    t = [(-v) % R for v in a]
    fixed = [[1 if r < USABLE else 0 for r in range(N)]]
    assert all(P.eval_expr(cs.gates[0][0], r, N, 1, fixed, [a, t], []) == 0 for r in range(USABLE))
    ev = P.Evaluator.new(cs)
    assert ("Negate", ("Intermediate", 0)) not in ev.calculations                 # -a is never computed
    assert HP.evaluator_parts(HP.ConstraintSystem.like(cs))["calculations"] == ev.calculations
    oparams = PR.Params(K, S_TOXIC)
    opk = PR.keygen(oparams, cs, fixed, [])
    proof = PR.create_proof(oparams, opk, [a, t], [], HP.SeededRng(1))
    assert not PR.verify_proof(oparams, opk.vk, [], proof)
    # the witness the lowered form accepts (t = +a) is not a witness of the gate as written either; with this one-gate
    # circuit its numerator is identically zero, so there is no h(X) piece to commit to
    with pytest.raises(PR.TranscriptError):
        PR.create_proof(oparams, opk, [a, list(a)], [], HP.SeededRng(1))


@pytest.mark.parametrize("two_classes", [False, True])
@pytest.mark.parametrize("seed", range(12))
def test_random_circuit_lowered_quotient_program(seed, two_classes, monkeypatch):
    """The C++ lowering of the whole evaluate_h program (b2_quotient_program_create: inlining, dead-code removal,
    depth-first scheduling, slot allocation, derived challenge powers; host-only code of the product) on the random
    circuits -- also with the shared-slot target lowered to 2, so that these small programs go through the two-class
    allocator (values with live ranges of 3+ instructions move to the global class until 2 shared slots remain):
    the dumped program, interpreted with big ints over the oracle's cosets, must equal the oracle's
    evaluate_h on every row of the extended domain."""
    from test_quotient_lowering import interpret
    if two_classes:
        monkeypatch.setenv("B2_Q_SHARED_TARGET", "2")
        monkeypatch.setenv("B2_Q_GLOBAL_MIN_LIVE", "3")
    cs, fixed, advice, instance, mapping = build(seed)
    rng = random.Random(1000 + seed)
    d = o.EvaluationDomain(cs.degree(), K)
    theta, beta, gamma, y = (rng.randrange(R) for _ in range(4))
    sigmas = P.permutation_sigmas(cs, d, mapping)
    perm_z = P.permutation_commit(cs, d, sigmas, advice, fixed, instance, beta, gamma, rng)
    lookups = []
    for lk in cs.lookups:
        input_sets, table, m = P.logup_compress(cs, d, lk, theta, advice, fixed, instance, rng)
        zs = [P.blind_to_n(z, N, rng) for z in P.logup_commit_z(cs, d, input_sets, table, m, beta)]
        lookups.append((zs, m))
    shuffle_z = [P.blind_to_n(P.shuffle_commit_product(cs, d, g, theta, beta, advice, fixed, instance), N, rng)
                 for g in cs.shuffles]
    ext = lambda col: d.coeff_to_extended(d.lagrange_to_coeff(col))                     # noqa: E731
    l0, l_last, l_active = P.lagrange_basis_cosets(cs, d)
    cz_fixed, cz_adv, cz_inst = [ext(c) for c in fixed], [ext(c) for c in advice], [ext(c) for c in instance]
    cz_sigma, cz_perm = [ext(c) for c in sigmas], [ext(c) for c in perm_z]
    cz_lookups = [{"z_cosets": [ext(z) for z in zs], "m_coset": ext(m)} for zs, m in lookups]
    cz_shuffles = [ext(z) for z in shuffle_z]
    ev = P.Evaluator.new(cs)
    want = P.evaluate_h(ev, cs, d, cz_fixed, cz_adv, cz_inst, l0, l_last, l_active, cz_sigma, y, beta, gamma, theta,
                        cz_lookups, cz_shuffles, cz_perm)
    hcs = HP.ConstraintSystem.like(cs)
    Ev = HP.build_evaluator(hcs)
    prog = Ev.program(len(perm_z), [len(zs) for zs, _ in lookups], len(shuffle_z))
    aux = [l0, l_last, l_active] + cz_sigma + cz_perm
    for lk in cz_lookups:
        aux += lk["z_cosets"] + [lk["m_coset"]]
    aux += cz_shuffles
    challenges = [beta, gamma, theta, y]
    dlt = beta * d.g_coset % R
    for _ in cs.permutation_columns:
        challenges.append(dlt)
        dlt = dlt * P.FR_DELTA % R
    rotations = list(ev.rotations)
    for r in (0, 1, -(BF + 1)):
        if r not in rotations:
            rotations.append(r)
    constants = list(ev.constants)
    got = interpret(prog, rotations, constants, cz_fixed + cz_adv + cz_inst + aux, challenges, d.extended_len(),
                    1 << (d.extended_k - d.k), 1, d.extended_omega)
    assert got == want
    info = prog.info()
    assert info["n_slots"] <= 24 and info["n_instr"] > 0
    assert info["n_slots"] == info["n_slots_shared"] + info["n_slots_global"]
    if two_classes:
        assert info["n_slots_global"] > 0 or info["n_slots_shared"] <= 2
    else:
        assert info["n_slots_global"] == 0 or info["n_slots"] >= 10
    if seed < 6:
        # the third restatement: oracle/cpu_ref.c's row loop (bench.py's CPU baseline for evaluate_h) on the flat program
        from oracle import cref
        f = Ev.flat_h_program(len(perm_z), [len(zs) for zs, _ in lookups], len(shuffle_z))
        c_out = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"],
                                   [enc(c) for c in cz_fixed], [enc(c) for c in cz_adv], [enc(c) for c in cz_inst],
                                   [enc(c) for c in aux], enc(challenges), d.extended_k, 1 << (d.extended_k - d.k),
                                   x0=enc([1])[0], step=enc([d.extended_omega])[0], threads=2)
        assert np.array_equal(c_out, enc(want))


def test_random_expressions_through_the_z_column_compiler():
    """grand_product.ExprCompiler + the C++ lowering on random expression lists (zeros and `0 - b` included: the
    compression path evaluates expressions as written, evaluate_with_theta has no quirk): the dumped program interpreted
    at rows = n equals evaluate_with_theta (plonk/evaluation.rs:2330-2398)"""
    from halo2_gpu_specific_b200.evaluation import QuotientProgram
    from halo2_gpu_specific_b200.grand_product import ExprCompiler
    from test_quotient_lowering import interpret
    rng = random.Random(77)
    fixed = [[rng.randrange(R) for _ in range(N)] for _ in range(N_FIXED_CONST)]
    advice = [[rng.randrange(R) for _ in range(N)] for _ in range(N_IN)]
    instance = [[rng.randrange(R) for _ in range(N)]]
    for trial in range(40):
        exprs = [random_expression(rng, 3, 4) for _ in range(rng.randint(1, 3))]
        if trial % 5 == 0:
            exprs.append(P.Sub(P.Const(0), P.Advice(1, 1)))
        if trial % 7 == 0:
            exprs.append(P.Instance(0, -1))
        theta = rng.randrange(R)
        c = ExprCompiler()
        res = c.emit(("Store", c.compress(exprs)))
        prog = QuotientProgram(c.rotations, c.constants, c.calcs, res, N_FIXED_CONST, N_IN, 1, 0, 3)
        got = interpret(prog, c.rotations, c.constants, fixed + advice + instance, [0, 0, theta], N, 1, 1, 1)
        prog.free()
        assert got == P.evaluate_with_theta(exprs, N, 1, fixed, advice, instance, theta), trial
