"""CPU-only: the host logic of the quotient engine -- the Python mirror's program builder
(halo2_gpu_specific_b200/evaluation.py) and the C++ lowering behind b2_quotient_program_create
(Store inlining, dead-code elimination, slot allocation, derived challenge powers) -- checked by
interpreting the DUMPED lowered program with big ints and comparing every extended-domain row with the
oracle's evaluate_h (oracle/plonk.py, restatement of plonk/evaluation.rs:778-1226)."""
import pytest

import plonk_fixture as fxm
from oracle import bn254 as o
from oracle import plonk as P

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import evaluation as E
from halo2_gpu_specific_b200._lib import B2Error

R = o.R_MOD


def make_evaluator(fx):
    ev, cs = fx["ev"], fx["cs"]
    return h2.Evaluator(ev.rotations, ev.constants, ev.calculations, ev.value_parts, ev.lookup_results,
                        ev.shuffle_results, cs.num_fixed, cs.num_advice, cs.num_instance, cs.permutation_columns,
                        cs.degree(), cs.blinding_factors())


def interpret(prog, rotations, constants, columns, challenges, rows, rot_scale, x0, x_step):
    """reference semantics of the lowered form (include/b2pcs.h, b2_quotient_program_dump)"""
    instr, result, derived = prog.dump()
    ch = list(challenges) + [pow(challenges[c], p, R) for c, p in derived]
    n_slots = prog.info()["n_slots"]
    out = []
    x = x0
    for row in range(rows):
        slots = [None] * n_slots

        def fetch(w):
            kind, rot, idx = w >> 28, (w >> 20) & 0xff, w & 0xfffff
            if kind == 0:
                return constants[idx]
            if kind == 1:
                assert slots[idx] is not None, "read of an unwritten slot"
                return slots[idx]
            if kind == 2:
                return columns[idx][(row + rotations[rot] * rot_scale) % rows]
            if kind == 3:
                return ch[idx]
            return x

        for ins in instr:
            op, dst, a, b = ins[:4]
            va = fetch(a)
            if op == 3:
                r = (-va) % R
            elif op == 4:
                r = va
            elif op in (5, 6):                         # fused a * b +- c * d
                vc, vd = fetch(ins[4]), fetch(ins[5])
                r = (va * fetch(b) + (vc * vd if op == 5 else -vc * vd)) % R
            else:
                vb = fetch(b)
                r = (va * vb) % R if op == 2 else ((va + vb) % R if op == 0 else (va - vb) % R)
            slots[dst] = r
        out.append(fetch(result))
        x = x * x_step % R
    return out


@pytest.fixture(scope="module")
def fx():
    return fxm.build(k=5, seed=11)


def test_lowered_program_matches_oracle(fx):
    Ev = make_evaluator(fx)
    cz = fxm.cosets(fx)
    want = fxm.oracle_h(fx, cz)
    prog = Ev.program(len(fx["perm_z"]), [len(l["z"]) for l in fx["lookups_lagrange"]], len(fx["shuffle_z"]))
    info = prog.info()
    assert info["n_slots"] < info["n_instr"] and info["n_mul"] > 0
    ops = [ins[0] for ins in prog.dump()[0]]
    assert ops.count(5) > 0 and ops.count(6) > 0, "the fixture must exercise both fused forms (a*b + c*d, a*b - c*d)"
    d = fx["domain"]
    aux = [cz["l0"], cz["l_last"], cz["l_active_row"]] + cz["sigma"] + cz["perm_z"]
    for lk in cz["lookups"]:
        aux += lk["z_cosets"] + [lk["m_coset"]]
    aux += cz["shuffles"]
    columns = cz["fixed"] + cz["advice"] + cz["instance"] + aux
    challenges = [fx["beta"], fx["gamma"], fx["theta"], fx["y"]]
    dlt = fx["beta"] * d.g_coset % R
    for _ in fx["cs"].permutation_columns:
        challenges.append(dlt)
        dlt = dlt * P.FR_DELTA % R
    # rotations / constants as the builder extended them: re-derive through a second build
    rotations = list(fx["ev"].rotations)
    for r in (0, 1, -(fx["cs"].blinding_factors() + 1)):
        if r not in rotations:
            rotations.append(r)
    constants = list(fx["ev"].constants)
    if 1 not in constants:
        constants.append(1)
    got = interpret(prog, rotations, constants, columns, challenges, d.extended_len(),
                    1 << (d.extended_k - d.k), 1, d.extended_omega)
    assert got == want


def test_gates_only_program(fx):
    """no permutation / lookups / shuffles: only the value_parts fold"""
    ev, cs = fx["ev"], fx["cs"]
    Ev = h2.Evaluator(ev.rotations, ev.constants, ev.calculations, ev.value_parts, [], [], cs.num_fixed,
                      cs.num_advice, cs.num_instance, [], cs.degree(), cs.blinding_factors())
    prog = Ev.program(0, [], 0)
    cz = fxm.cosets(fx)
    d = fx["domain"]
    cs2 = P.ConstraintSystem(cs.num_fixed, cs.num_advice, cs.num_instance, cs.degree(), cs.blinding_factors())
    cs2.gates = cs.gates
    ev2 = P.Evaluator.new(cs2)
    want = P.evaluate_h(ev2, cs2, d, cz["fixed"], cz["advice"], cz["instance"], cz["l0"], cz["l_last"],
                        cz["l_active_row"], [], fx["y"], fx["beta"], fx["gamma"], fx["theta"], [], [], [])
    rotations = list(ev.rotations)
    for r in (0, 1, -(cs.blinding_factors() + 1)):
        if r not in rotations:
            rotations.append(r)
    columns = cz["fixed"] + cz["advice"] + cz["instance"] + [cz["l0"], cz["l_last"], cz["l_active_row"]]
    got = interpret(prog, rotations, list(ev.constants), columns, [fx["beta"], fx["gamma"], fx["theta"], fx["y"]],
                    d.extended_len(), 1 << (d.extended_k - d.k), 1, d.extended_omega)
    assert got == want
    # the lookups' calculations are dead code here and must have been eliminated
    full = make_evaluator(fx).program(len(fx["perm_z"]), [len(l["z"]) for l in fx["lookups_lagrange"]], 1).info()
    assert prog.info()["n_instr"] < full["n_instr"] // 2


def test_lowering_rejects_bad_programs():
    mk = lambda calcs, result, **kw: E.QuotientProgram([0], [0, 1], calcs, result, kw.get("nf", 1), 1, 0, 0, 4)  # noqa: E731
    with pytest.raises(B2Error):   # forward reference
        mk([("Add", ("Intermediate", 1), ("Constant", 0)), ("Store", ("Constant", 1))], ("Intermediate", 0))
    with pytest.raises(B2Error):   # column out of range
        mk([("Store", ("Fixed", 3, 0))], ("Intermediate", 0))
    with pytest.raises(B2Error):   # rotation index out of range
        mk([("Store", ("Advice", 0, 2))], ("Intermediate", 0))
    with pytest.raises(B2Error):   # constant out of range
        mk([("Store", ("Constant", 9))], ("Intermediate", 0))
    p = mk([("Store", ("Fixed", 0, 0))], ("Intermediate", 0))   # a pure column read needs no instruction
    assert p.info()["n_instr"] == 0
    p.free()


def test_lc_challenge_powers_are_derived():
    calcs = [("Store", ("Advice", 0, 0)),
             ("LcChallenge", ("Intermediate", 0), ("Intermediate", 0), "Beta", 3),
             ("LcChallenge", ("Intermediate", 1), ("Intermediate", 0), "Beta", 3),
             ("LcChallenge", ("Intermediate", 2), ("Intermediate", 0), "Gamma", 1)]
    p = E.QuotientProgram([0], [0, 1], calcs, ("Intermediate", 3), 0, 1, 0, 0, 4)
    instr, result, derived = p.dump()
    assert derived == [(0, 3)]            # beta^3 once; gamma^1 is gamma itself (evaluation.rs:208-211)
    a0 = [5, 7, 11, 13]
    got = interpret(p, [0], [0, 1], [a0], [2, 3, 0, 0], 4, 1, 1, 1)
    want = []
    for v in a0:
        t = (v + 8) * v % R
        t = (t + 8) * v % R
        want.append((t + 3) * v % R)
    assert got == want


def test_c_restatement_matches_python_oracle(fx):
    """oracle/cpu_ref.c ref_quotient_eval (Calculation::evaluate over all rows) on the flat evaluate_h program
    == oracle/plonk.py evaluate_h: two restatements of plonk/evaluation.rs agree."""
    import numpy as np
    from oracle import cref
    Ev = make_evaluator(fx)
    cz = fxm.cosets(fx)
    want = fxm.oracle_h(fx, cz)
    f = Ev.flat_h_program(len(fx["perm_z"]), [len(l["z"]) for l in fx["lookups_lagrange"]], len(fx["shuffle_z"]))
    d = fx["domain"]
    enc = o.fr_encode
    aux = [cz["l0"], cz["l_last"], cz["l_active_row"]] + cz["sigma"] + cz["perm_z"]
    for lk in cz["lookups"]:
        aux += lk["z_cosets"] + [lk["m_coset"]]
    aux += cz["shuffles"]
    assert len(aux) == f["n_aux"]
    challenges = [fx["beta"], fx["gamma"], fx["theta"], fx["y"]]
    dlt = fx["beta"] * d.g_coset % R
    for _ in fx["cs"].permutation_columns:
        challenges.append(dlt)
        dlt = dlt * P.FR_DELTA % R
    got = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"],
                             [enc(c) for c in cz["fixed"]], [enc(c) for c in cz["advice"]],
                             [enc(c) for c in cz["instance"]], [enc(c) for c in aux], enc(challenges),
                             d.extended_k, 1 << (d.extended_k - d.k), x0=enc([1])[0], step=enc([d.extended_omega])[0],
                             threads=3)
    assert np.array_equal(got, enc(want))


def test_long_lived_values_move_to_the_global_slot_class(monkeypatch):
    """A program whose gates share sub-expressions far apart (tools/quotient_bench.py long_lived: the first gates'
    values are read again by extra gates at the end of the list, so the y-fold keeps them alive across the whole
    program): with every slot in shared memory the live width costs resident CTAs; the lowering moves the
    longest-lived values to the global class (from 10 live values on) until at most 7 shared slots remain (b2_quotient_program_slot_classes).
    Both lowerings are interpreted with big ints on every row and must agree with each other and with the C
    restatement of Calculation::evaluate on the flat program."""
    import os
    import random
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import quotient_bench as qb
    from oracle import cref
    ev, lookups, shuffles, n_sets = qb.synthetic_evaluator(A=12, F=6, gates=24, lookups=(2, 1), shuffles=1, perm_cols=6,
                                                           long_lived=14)
    monkeypatch.setenv("B2_Q_HYBRID", "0")
    plain = ev.program(n_sets, list(lookups), shuffles)
    ev2, _, _, _ = qb.synthetic_evaluator(A=12, F=6, gates=24, lookups=(2, 1), shuffles=1, perm_cols=6, long_lived=14)
    monkeypatch.delenv("B2_Q_HYBRID")
    hybrid = ev2.program(n_sets, list(lookups), shuffles)
    ip, ih = plain.info(), hybrid.info()
    assert ip["n_slots_global"] == 0 and ip["n_slots_shared"] == ip["n_slots"] > 7 + 6
    assert ih["n_slots_shared"] <= 7 and ih["n_slots_global"] >= 6
    assert ih["n_slots"] == ih["n_slots_shared"] + ih["n_slots_global"]
    assert (ih["n_instr"], ih["n_mul"], ih["n_addsub"]) == (ip["n_instr"], ip["n_mul"], ip["n_addsub"])
    # a program without such values is left alone, and so is one of fewer than 10 live values (measured: 8 values are
    # faster all in shared memory at 6 CTAs per SM than as 7 + 1)
    ev3, lk3, sh3, ns3 = qb.synthetic_evaluator(A=12, F=6, gates=24, lookups=(2, 1), shuffles=1, perm_cols=6)
    assert ev3.program(ns3, list(lk3), sh3).info()["n_slots_global"] == 0
    for ll, classes in ((6, (8, 0)), (7, (9, 0)), (8, (7, 3))):
        e4, lk4, sh4, ns4 = qb.synthetic_evaluator(gates=96, long_lived=ll)
        i4 = e4.program(ns4, list(lk4), sh4).info()
        assert (i4["n_slots_shared"], i4["n_slots_global"]) == classes
    log_rows, rot_scale = 4, 2
    rows = 1 << log_rows
    ncols = plain.n_fixed + plain.n_advice + plain.n_instance + plain.n_aux
    cols_m = cref.random_fr_mont(rows * ncols, 0xB20000AA).reshape(ncols, rows, 4)
    cols = [o.fr_decode(c) for c in cols_m]
    rng = random.Random(8)
    challenges = [rng.randrange(R) for _ in range(plain.n_challenges)]
    x0, step = rng.randrange(R), rng.randrange(R)
    f = ev.flat_h_program(n_sets, list(lookups), shuffles)
    a = interpret(plain, f["rotations"], f["constants"], cols, challenges, rows, rot_scale, x0, step)
    b = interpret(hybrid, f["rotations"], f["constants"], cols, challenges, rows, rot_scale, x0, step)
    assert a == b
    nf, na, ni = plain.n_fixed, plain.n_advice, plain.n_instance
    enc = o.fr_encode
    want = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"], list(cols_m[:nf]),
                              list(cols_m[nf:nf + na]), list(cols_m[nf + na:nf + na + ni]), list(cols_m[nf + na + ni:]),
                              enc(challenges), log_rows, rot_scale, x0=enc([x0])[0], step=enc([step])[0], threads=2)
    assert np.array_equal(enc(b), want)
    # the global class holds the long-lived values: every global slot is written once per row and read later
    instr, _, _ = hybrid.dump()
    writes = {}
    for pos, ins in enumerate(instr):
        if ins[1] >= ih["n_slots_shared"]:
            writes.setdefault(ins[1], []).append(pos)
    assert len(writes) == ih["n_slots_global"]
    plain.free(); hybrid.free()


def test_gate_order_of_the_proof_circuit_and_the_slot_classes():
    """the stand-in circuit of the k = 22 proof (tools/zkwasm_shape_circuit.py) in both gate orders: with each extra
    gate next to the gate whose factor it shares the lowered evaluate_h program needs 5 live values; with the extra
    gates after all product gates -- the same constraints -- every shared factor stays alive across the y-fold, 17
    values, of which the lowering keeps 7 in shared memory and 10 in the global class.  Same instruction and product
    counts either way: the gate order is the front-end's choice and must not cost the engine its resident CTAs."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import zkwasm_shape_circuit as zk
    from halo2_gpu_specific_b200 import plonk as HP
    infos = []
    for at_end in (False, True):
        cs = HP.ConstraintSystem(**zk.constraint_system_args(extra_gates=300, extras_at_end=at_end))
        n_perm = (len(cs.permutation_columns) + cs.degree() - 3) // (cs.degree() - 2)
        prog = HP.build_evaluator(cs).program(n_perm, [len(lk["input_expressions_sets"]) for lk in cs.lookups],
                                              len(cs.shuffles))
        infos.append(prog.info())
        prog.free()
    near, far = infos
    assert (near["n_slots_shared"], near["n_slots_global"]) == (5, 0)
    assert (far["n_slots_shared"], far["n_slots_global"]) == (7, 10)
    assert (near["n_instr"], near["n_mul"], near["n_addsub"]) == (far["n_instr"], far["n_mul"], far["n_addsub"])
