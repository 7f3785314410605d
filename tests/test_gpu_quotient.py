"""GPU parity of the quotient engine (b2_quotient_* through the Python mirror of Evaluator::evaluate_h)
against the oracle's evaluate_h (oracle/plonk.py), bit-exact on every extended-domain row."""
import os
import random

import numpy as np
import pytest

import plonk_fixture as fxm
from oracle import bn254 as o
from oracle import plonk as P

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import evaluation as E
from test_quotient_lowering import interpret, make_evaluator

pytestmark = pytest.mark.gpu
R = o.R_MOD


def enc(v):
    return o.fr_encode(v)


def run_gpu(fx, Ev, dom, to_coeff=False, **kw):
    c = fx["coeff"]
    l0, l_last, l_active = P.lagrange_basis_cosets(fx["cs"], fx["domain"])
    lookups = [{"z": [enc(z) for z in zs], "m": enc(m)} for zs, m in zip(c["lookup_z"], c["lookup_m"])]
    return Ev.evaluate_h(dom, [enc(p) for p in c["fixed"]], [enc(p) for p in c["advice"]],
                         [enc(p) for p in c["instance"]], enc(l0), enc(l_last), enc(l_active),
                         [enc(p) for p in c["sigma"]], fx["y"], fx["beta"], fx["gamma"], fx["theta"], lookups,
                         [enc(p) for p in c["shuffle_z"]], [enc(p) for p in c["perm_z"]], to_coeff=to_coeff, **kw)


@pytest.mark.parametrize("mode", ["extended", "cosets"])
@pytest.mark.parametrize("k,j", [(5, None), (5, 9), (6, None)])
def test_evaluate_h_matches_oracle(gpu, k, j, mode):
    fx = fxm.build(k=k, seed=3 + k, domain_j=j)
    dom = h2.EvaluationDomain(j or fx["cs"].degree(), k)
    got = run_gpu(fx, make_evaluator(fx), dom, mode=mode)
    want = fxm.oracle_h(fx)
    assert np.array_equal(got, enc(want))


def test_coset_subset_leaves_other_rows(gpu):
    """the multi-GPU split: a rank evaluates only its cosets; together they tile the extended domain"""
    fx = fxm.build(k=5, seed=31)
    dom = h2.EvaluationDomain(fx["cs"].degree(), 5)
    want = enc(fxm.oracle_h(fx))
    Ev = make_evaluator(fx)
    nc = 1 << (dom.extended_k - dom.k)
    merged = np.zeros_like(want)
    for rank in range(2):
        mine = list(range(rank, nc, 2))
        part = run_gpu(fx, Ev, dom, cosets=mine)
        for c in mine:
            merged[c::nc] = part[c::nc]
    assert np.array_equal(merged, want)


def test_coset_transform_matches_extended(gpu):
    """b2_ntt_desc.coset_gen: coset c of coeff_to_extended's output without zero padding"""
    from oracle import cref
    k = 10
    dom = h2.EvaluationDomain(5, k)
    a = cref.random_fr_mont(3 << k, 0xB2000055).reshape(3, 1 << k, 4)
    ext = dom.coeff_to_extended(a)
    nc = 1 << (dom.extended_k - k)
    src = E.DeviceBuffer(3 << k).upload(a)
    dst = E.DeviceBuffer(3 << k)
    for c in range(nc):
        g = dom._zeta * pow(dom._ext_omega, c, R) % R
        E.coeff_to_coset_dev(dom, src.ptr, 3, g, dst.ptr)
        got = dst.download().reshape(3, 1 << k, 4)
        assert np.array_equal(got, ext[:, c::nc])
    src.free(); dst.free()


def test_h_coefficients_match_oracle(gpu):
    """vanishing::Argument::construct's h(X): numerator / (X^n - 1), extended_to_coeff (vanishing/prover.rs:64-96)"""
    fx = fxm.build(k=5, seed=21)
    d = fx["domain"]
    dom = h2.EvaluationDomain(fx["cs"].degree(), 5)
    want = enc(d.extended_to_coeff(d.divide_by_vanishing_poly(fxm.oracle_h(fx))))
    for mode in ("extended", "cosets"):
        got = run_gpu(fx, make_evaluator(fx), dom, to_coeff=True, mode=mode)
        assert np.array_equal(got, want), mode


def test_spilled_slots_match(gpu, monkeypatch):
    fx = fxm.build(k=5, seed=5)
    dom = h2.EvaluationDomain(fx["cs"].degree(), 5)
    a = run_gpu(fx, make_evaluator(fx), dom)
    monkeypatch.setenv("B2_Q_FORCE_SPILL", "1")
    b = run_gpu(fx, make_evaluator(fx), dom)
    assert np.array_equal(a, b)
    assert np.array_equal(a, enc(fxm.oracle_h(fx)))


def test_large_random_program_spot_rows(gpu):
    """2^18 rows, 12 columns, a random straight-line program with rotations and the coset point: 48 random
    rows are recomputed with big ints from the dumped program."""
    rng = random.Random(99)
    log_rows, ncol = 18, 12
    rows = 1 << log_rows
    rotations = [0, 1, -1, 5, -7]
    constants = [0, 1] + [rng.randrange(R) for _ in range(6)]
    calcs = []
    for i in range(120):
        def src():
            t = rng.random()
            if t < 0.45 and calcs:
                return ("Intermediate", rng.randrange(len(calcs)))
            if t < 0.8:
                return ("Advice", rng.randrange(ncol), rng.randrange(len(rotations)))
            if t < 0.9:
                return ("Constant", rng.randrange(len(constants)))
            if t < 0.95:
                return ("Challenge", rng.randrange(4))
            return ("CosetX",)
        op = rng.choice(["Add", "Sub", "Mul", "Mul", "Negate", "LcTheta", "AddChallenge", "LcChallenge", "Store"])
        if op in ("Add", "Sub", "Mul", "LcTheta"):
            calcs.append((op, src(), src()))
        elif op in ("Negate", "Store"):
            calcs.append((op, src()))
        elif op == "AddChallenge":
            calcs.append((op, src(), rng.choice(["Beta", "Gamma"])))
        else:
            calcs.append((op, src(), src(), rng.choice(["Beta", "Gamma"]), rng.randrange(1, 5)))
    # fold the last 30 calculations so that most of the program is live
    acc = ("Intermediate", len(calcs) - 30)
    for i in range(len(calcs) - 29, len(calcs) - 0):
        calcs.append(("MulChAdd", acc, ("Intermediate", i), E.CH_Y))
        acc = ("Intermediate", len(calcs) - 1)
    prog = E.QuotientProgram(rotations, constants, calcs, acc, 0, ncol, 0, 0, 4)
    from oracle import cref
    cols = cref.random_fr_mont(rows * ncol, 0xB2000077).reshape(ncol, rows, 4)
    buf = E.DeviceBuffer(rows * ncol).upload(cols)
    out = E.DeviceBuffer(rows)
    challenges = [rng.randrange(R) for _ in range(4)]
    x0, step = rng.randrange(R), rng.randrange(R)
    prog.eval(log_rows, 4, [], [buf.ptr + c * rows * 32 for c in range(ncol)], [], [], challenges, out.ptr,
              x0=x0, x_step=step)
    got = out.download()
    buf.free(); out.free()
    instr, result, derived = prog.dump()
    ch = challenges + [pow(challenges[c], p, R) for c, p in derived]
    picks = [0, 1, rows - 1, rows - 2] + [rng.randrange(rows) for _ in range(44)]
    for row in picks:
        slots = {}

        def fetch(w):
            kind, rot, idx = w >> 28, (w >> 20) & 0xff, w & 0xfffff
            if kind == 0:
                return constants[idx]
            if kind == 1:
                return slots[idx]
            if kind == 2:
                return o.fr_decode(cols[idx][(row + rotations[rot] * 4) % rows][None])[0]
            if kind == 3:
                return ch[idx]
            return x0 * pow(step, row, R) % R

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
        assert o.fr_decode(got[row][None])[0] == fetch(result), f"row {row}"
    prog.free()


def test_zkwasm_shape_program_full_compare(gpu):
    """the benchmark's synthetic zkWasm-scale program (tools/quotient_bench.py: 1200+ instructions, 156 columns)
    at 2^13 rows with rot_scale 4: every row against the C restatement of Calculation::evaluate."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import quotient_bench as qb
    from oracle import cref
    ev, lookups, shuffles, n_sets = qb.synthetic_evaluator(gates=40)
    f = ev.flat_h_program(n_sets, lookups, shuffles)
    prog = ev.program(n_sets, lookups, shuffles)
    log_rows = 13
    rows = 1 << log_rows
    ncols = prog.n_fixed + prog.n_advice + prog.n_instance + prog.n_aux
    cols = cref.random_fr_mont(rows * ncols, 0xB2000088).reshape(ncols, rows, 4)
    rng = random.Random(5)
    challenges = [rng.randrange(R) for _ in range(prog.n_challenges)]
    x0, step = rng.randrange(R), rng.randrange(R)
    scale = cref.random_fr_mont(4, 0xB2000089)
    buf = E.DeviceBuffer(rows * ncols).upload(cols)
    out = E.DeviceBuffer(rows)
    ptrs = [buf.ptr + c * rows * 32 for c in range(ncols)]
    nf, na, ni = prog.n_fixed, prog.n_advice, prog.n_instance
    prog.eval(log_rows, 4, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
              out.ptr, x0=x0, x_step=step)
    got = out.download()
    want = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"], list(cols[:nf]),
                              list(cols[nf:nf + na]), list(cols[nf + na:nf + na + ni]), list(cols[nf + na + ni:]),
                              enc(challenges), log_rows, 4, x0=enc([x0])[0], step=enc([step])[0], threads=4)
    assert np.array_equal(got, want)
    # the same with the vanishing scale folded into the store
    prog.eval(log_rows, 4, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
              out.ptr, x0=x0, x_step=step, scale=scale)
    got = out.download()
    sc = o.fr_decode(scale)
    wd = o.fr_decode(want[:64])
    gd = o.fr_decode(got[:64])
    assert gd == [w * sc[i % 4] % R for i, w in enumerate(wd)]
    buf.free(); out.free()


@pytest.mark.parametrize("log_rows,row_range", [(18, None), (13, None), (18, (5000, 200000))])
def test_global_slot_class_matches_c_restatement(gpu, monkeypatch, log_rows, row_range):
    """a program with long-lived shared values (tools/quotient_bench.py long_lived): the lowering keeps 7 slots in shared
    memory and the rest in the per-CTA global scratch (quotient_eval_kernel<true, true>, one wave of resident CTAs
    striding over the rows: 2^18 rows are several trips per CTA, 2^13 less than one wave); every row against the C
    restatement of Calculation::evaluate, and against the same circuit lowered with every slot in shared memory"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import quotient_bench as qb
    from oracle import cref
    shape = dict(A=12, F=6, gates=24, lookups=(2, 1), shuffles=1, perm_cols=6, long_lived=14)
    ev, lookups, shuffles, n_sets = qb.synthetic_evaluator(**shape)
    f = ev.flat_h_program(n_sets, list(lookups), shuffles)
    prog = ev.program(n_sets, list(lookups), shuffles)
    info = prog.info()
    assert info["n_slots_shared"] <= 7 and info["n_slots_global"] >= 6
    monkeypatch.setenv("B2_Q_HYBRID", "0")
    plain = qb.synthetic_evaluator(**shape)[0].program(n_sets, list(lookups), shuffles)
    monkeypatch.delenv("B2_Q_HYBRID")
    assert plain.info()["n_slots_global"] == 0
    rows = 1 << log_rows
    ncols = prog.n_fixed + prog.n_advice + prog.n_instance + prog.n_aux
    cols = cref.random_fr_mont(rows * ncols, 0xB20000BB).reshape(ncols, rows, 4)
    rng = random.Random(6)
    challenges = [rng.randrange(R) for _ in range(prog.n_challenges)]
    x0, step = rng.randrange(R), rng.randrange(R)
    buf = E.DeviceBuffer(rows * ncols).upload(cols)
    out, out2 = E.DeviceBuffer(rows), E.DeviceBuffer(rows)
    ptrs = [buf.ptr + c * rows * 32 for c in range(ncols)]
    nf, na, ni = prog.n_fixed, prog.n_advice, prog.n_instance
    kw = {} if row_range is None else dict(row_begin=row_range[0], row_count=row_range[1])
    for p, dst in ((prog, out), (plain, out2)):
        p.eval(log_rows, 4, ptrs[:nf], ptrs[nf:nf + na], ptrs[nf + na:nf + na + ni], ptrs[nf + na + ni:], challenges,
               dst.ptr, x0=x0, x_step=step, **kw)
    got, got_plain = out.download(), out2.download()
    want = cref.quotient_eval(f["rotations"], enc(f["constants"]), f["calcs"], f["result"], list(cols[:nf]),
                              list(cols[nf:nf + na]), list(cols[nf + na:nf + na + ni]), list(cols[nf + na + ni:]),
                              enc(challenges), log_rows, 4, x0=enc([x0])[0], step=enc([step])[0], threads=8)
    if row_range is not None:          # the result of row i lands at out[i - row_begin]
        want = want[row_range[0]:row_range[0] + row_range[1]]
        got, got_plain = got[:row_range[1]], got_plain[:row_range[1]]
    assert np.array_equal(got, want)
    assert np.array_equal(got_plain, want)
    buf.free(); out.free(); out2.free()
    prog.free(); plain.free()


def test_sharded_path_single_rank(gpu):
    """parallel.sharded_evaluate_h with one rank: compact task output, interleave, extended_to_coeff"""
    from halo2_gpu_specific_b200 import parallel
    fx = fxm.build(k=5, seed=41)
    d = fx["domain"]
    dom = h2.EvaluationDomain(fx["cs"].degree(), 5)
    c = fx["coeff"]
    l0, l_last, l_active = P.lagrange_basis_cosets(fx["cs"], d)
    lookups = [{"z": [enc(z) for z in zs], "m": enc(m)} for zs, m in zip(c["lookup_z"], c["lookup_m"])]
    got = parallel.sharded_evaluate_h(make_evaluator(fx), dom, [enc(p) for p in c["fixed"]], [enc(p) for p in c["advice"]],
                                      [enc(p) for p in c["instance"]], enc(l0), enc(l_last), enc(l_active),
                                      [enc(p) for p in c["sigma"]], fx["y"], fx["beta"], fx["gamma"], fx["theta"], lookups,
                                      [enc(p) for p in c["shuffle_z"]], [enc(p) for p in c["perm_z"]])
    want = enc(d.extended_to_coeff(d.divide_by_vanishing_poly(fxm.oracle_h(fx))))
    assert np.array_equal(got, want)


def test_row_ranges_of_one_coset(gpu):
    """two ranks sharing a coset: each evaluates half of its rows"""
    fx = fxm.build(k=5, seed=43)
    dom = h2.EvaluationDomain(fx["cs"].degree(), 5)
    want = enc(fxm.oracle_h(fx))
    Ev = make_evaluator(fx)
    c = fx["coeff"]
    l0, l_last, l_active = P.lagrange_basis_cosets(fx["cs"], fx["domain"])
    lookups = [{"z": [enc(z) for z in zs], "m": enc(m)} for zs, m in zip(c["lookup_z"], c["lookup_m"])]
    prog = Ev.program(len(c["perm_z"]), [len(l["z"]) for l in lookups], len(c["shuffle_z"]))
    aux = [enc(p) for p in c["sigma"]] + [enc(p) for p in c["perm_z"]]
    for lk in lookups:
        aux += lk["z"] + [lk["m"]]
    aux += [enc(p) for p in c["shuffle_z"]]
    groups = [[enc(p) for p in c["fixed"]], [enc(p) for p in c["advice"]], [enc(p) for p in c["instance"]], aux]
    lag = [enc(v) for v in (l0, l_last, l_active)]
    ch = [fx["beta"], fx["gamma"], fx["theta"], fx["y"]]
    dlt = fx["beta"] * dom._zeta % R
    for _ in fx["cs"].permutation_columns:
        ch.append(dlt)
        dlt = dlt * P.FR_DELTA % R
    n, nc = dom.n, 4
    from halo2_gpu_specific_b200 import parallel
    out = E.DeviceBuffer(nc * n)
    for rank in range(8):
        tasks = parallel.quotient_tasks(nc, n, 8, rank)
        part = E.DeviceBuffer(sum(t[2] for t in tasks))
        Ev.evaluate_h_tasks(dom, groups, lag, ch, prog, tasks, part.ptr, compact=True, scaled=False, zeta=dom._zeta)
        got = part.download()
        pos = 0
        for cst, begin, count in tasks:
            rows = [nc * (begin + i) + cst for i in range(count)]
            assert np.array_equal(got[pos:pos + count], want[rows]), (rank, cst, begin)
            pos += count
        part.free()
    out.free()
