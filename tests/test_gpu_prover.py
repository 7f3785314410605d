"""GPU parity of the whole prover through the engine (halo2_gpu_specific_b200.plonk.create_proof over the C ABI):
proof bytes identical to the CPU oracle's (oracle/prover.py) under the same RNG, and acceptance by the oracle's
verify_proof (incl. the real pairing check) -- the reference's own prove -> verify test strategy -- up to the
benches/plonk.rs circuit at k = 18 (BASELINE config 4)."""
import os
import sys

import numpy as np
import pytest

import plonk_fixture as fxm
from oracle import bn254 as o
from oracle import plonk as P
from oracle import prover as PR

import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import grand_product as gp
from halo2_gpu_specific_b200 import plonk as HP

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import plonk_bench_circuit as bench_circuit  # noqa: E402

pytestmark = pytest.mark.gpu
R = o.R_MOD
enc, dec = o.fr_encode, o.fr_decode
S_TOXIC = 0x2B200B200B200B200B200B200B200B2001


def engine_side(k, oparams, cs, fixed, mapping, transcript_repr):
    params = h2.Params(k, oparams.g, oparams.g_lagrange)
    pk = HP.keygen(params, cs, fixed, mapping, transcript_repr=transcript_repr)
    return params, pk


@pytest.mark.parametrize("engine_kind", ["resident", "host_api"])
@pytest.mark.parametrize("k,seed,rng_seed", [(5, 11, 1), (6, 17, 2)])
def test_proof_bytes_match_oracle(gpu, k, seed, rng_seed, engine_kind):
    fx = fxm.build(k=k, seed=seed)
    ocs = fx["cs"]
    oparams = PR.Params(k, S_TOXIC)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)
    params, pk = engine_side(k, oparams, cs, np.stack([enc(c) for c in fx["fixed"]]),
                             np.array(fx["mapping"], dtype=np.int64), opk.vk.transcript_repr)
    try:
        # keygen on the device == keygen in the oracle
        assert pk.vk.fixed_commitments == opk.vk.fixed_commitments
        assert pk.vk.permutation_commitments == opk.vk.permutation_commitments
        for got, want in ((pk.sigmas, opk.sigmas), (pk.sigma_polys, opk.sigma_polys), (pk.fixed_polys, opk.fixed_polys)):
            for a, b in zip(got, want):
                assert np.array_equal(a, enc(b))
        assert np.array_equal(pk.l0, enc(opk.l0)) and np.array_equal(pk.l_last, enc(opk.l_last))
        assert np.array_equal(pk.l_active_row, enc(opk.l_active_row))
        inst = [fx["instance"][0][:4]]
        want = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(rng_seed))
        adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
        launches0 = gpu.lib().b2_launch_count(0)
        if engine_kind == "resident":
            eng = HP.ResidentEngine(params, pk.vk.domain)
            try:
                got = HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(rng_seed), engine=eng)
                # a second proof on the same engine reuses the resident proving key (and its cosets)
                again = HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(rng_seed), engine=eng)
                assert again == got
            finally:
                eng.free()
            assert HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(rng_seed)) == got     # default engine
        else:
            got = HP.create_proof(params, pk, adv, inst, HP.SeededRng(rng_seed), engine=HP.Engine(params, pk.vk.domain))
        assert gpu.lib().b2_launch_count(0) > launches0
        assert got == want
        assert PR.verify_proof(oparams, opk.vk, inst, got, pairing=(k == 5))
    finally:
        params.free()


def test_compress_expressions(gpu):
    fx = fxm.build(k=6, seed=17)
    cs, n = fx["cs"], fx["n"]
    dom = h2.EvaluationDomain(cs.degree(), fx["k"])
    adv, fixed, inst = [enc(c) for c in fx["advice"]], [enc(c) for c in fx["fixed"]], [enc(c) for c in fx["instance"]]
    lists = [[P.Advice(5)], [P.Fixed(4), P.Fixed(5)], [P.Const(3)], [P.Sum(P.Advice(1, 1), P.Const(2)), P.Instance(0, -1)]]
    got = gp.compress_expressions(dom, lists, adv, fixed, inst, fx["theta"])
    for g, ex in zip(got, lists):
        want = P.evaluate_with_theta(ex, n, 1, fx["fixed"], fx["advice"], fx["instance"], fx["theta"])
        assert np.array_equal(g, enc(want))


def test_commit_batch_against_g(gpu):
    """Params::commit for several polynomials shorter than n (the multiopen witnesses have n - 1 coefficients)"""
    k = 7
    oparams = PR.Params(k, S_TOXIC)
    params = h2.Params(k, oparams.g, oparams.g_lagrange)
    try:
        from oracle import cref
        x = cref.random_fr_mont(3 * 127, 0xB2000081).reshape(3, 127, 4)
        got = HP._points(params.commit_batch(x))
        assert got == [oparams.commit(dec(x[i])) for i in range(3)]
    finally:
        params.free()


@pytest.mark.parametrize("k", [1, 6, 11, 15])
def test_unsafe_setup_matches_oracle(gpu, k):
    """Params::unsafe_setup on the device (poly/commitment.rs:56-124): every point of g and g_lagrange"""
    oparams = PR.Params(k, S_TOXIC)
    params = h2.Params.unsafe_setup(k, S_TOXIC, precompute=False)
    try:
        assert np.array_equal(params.g.read(), oparams.g)
        assert np.array_equal(params.g_lagrange.read(), oparams.g_lagrange)
    finally:
        params.free()


def _bench_circuit(k):
    cs = HP.ConstraintSystem(**bench_circuit.constraint_system_args())
    fixed, advice, mapping = bench_circuit.build(k)
    ocs = P.ConstraintSystem(4, 3, 0, degree=5, blinding_factors=5)
    ocs.gates, ocs.permutation_columns = cs.gates, cs.permutation_columns
    ocs.advice_queries, ocs.fixed_queries, ocs.instance_queries = cs.advice_queries, cs.fixed_queries, cs.instance_queries
    return cs, ocs, fixed, advice, mapping


@pytest.mark.parametrize("k,pairing,engine_kind", [(8, True, "resident"), (14, False, "host_api"), (18, True, "resident"),
                                                   (20, False, "resident")])
def test_benches_plonk_circuit_proof_verifies(gpu, k, pairing, engine_kind):
    """BASELINE config 4: the benches/plonk.rs circuit, full create_proof on the engine, accepted by the oracle
    verifier.  The SRS is built on the device (Params.unsafe_setup) and the verifying key the oracle reads is the
    one the ENGINE's keygen produced (commitments computed on the device); the oracle side only needs the SRS
    trapdoor / [s]G2 and the constraint system."""
    cs, ocs, fixed, advice, mapping = _bench_circuit(k)
    oparams = PR.Params(k, S_TOXIC) if k <= 8 else PR.ParamsVerifier(k, S_TOXIC)
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        timings = {}
        eng = HP.ResidentEngine(params, pk.vk.domain) if engine_kind == "resident" else HP.Engine(params, pk.vk.domain)
        HP.create_proof(params, pk, advice.copy(), [], HP.SeededRng(0), engine=eng)              # warm-up
        proof = HP.create_proof(params, pk, advice.copy(), [], HP.SeededRng(k), timings=timings, engine=eng)
        assert len(proof) == 32 * ((3 + 1 + 1 + 4) + (3 + 4 + 1 + 3 + 2) + 2)
        ovk = PR.VerifyingKey(ocs, o.EvaluationDomain(cs.degree(), k), pk.vk.fixed_commitments,
                              pk.vk.permutation_commitments, pk.vk.transcript_repr)
        assert PR.verify_proof(oparams, ovk, [], proof, pairing=pairing)
        if k <= 8:
            # same bytes as the oracle prover, which also did its own keygen
            opk = PR.keygen(oparams, ocs, [dec(c) for c in fixed], [[(int(c), int(r)) for c, r in col] for col in mapping],
                            transcript_repr=pk.vk.transcript_repr)
            assert opk.vk.fixed_commitments == pk.vk.fixed_commitments
            assert PR.create_proof(oparams, opk, [dec(c) for c in advice], [], HP.SeededRng(k)) == proof
        # a broken witness (one product row off by one) must not verify
        bad = advice.copy()
        bad[2, 4] = enc([5])[0]
        proof_bad = HP.create_proof(params, pk, bad, [], HP.SeededRng(k), engine=eng)
        assert not PR.verify_proof(oparams, ovk, [], proof_bad)
        eng.free()
        print(f"k={k} {engine_kind} create_proof {sum(timings.values()):.4f} s: "
              + ", ".join(f"{a} {b:.4f}" for a, b in timings.items()))
    finally:
        params.free()


def _zk_shape(k, extra_gates, to_mont):
    import zkwasm_shape_circuit as zk
    args = zk.constraint_system_args(extra_gates=extra_gates)
    cs = HP.ConstraintSystem(**args)
    ocs = P.ConstraintSystem(args["num_fixed"], args["num_advice"], args["num_instance"], degree=5, blinding_factors=5)
    ocs.gates, ocs.lookups, ocs.shuffles = cs.gates, cs.lookups, cs.shuffles
    ocs.permutation_columns = cs.permutation_columns
    fixed, advice, public, mapping = zk.build(k, to_mont, seed=k)
    return cs, ocs, fixed, advice, public, mapping


@pytest.mark.parametrize("k", [7, 14])
def test_zkwasm_shaped_circuit(gpu, k):
    """BASELINE config 5's shape with a real witness: 64 advice, 32 fixed, 8 lookups (12 input sets, multiplicities
    counted on the device), 4 shuffles, 24 permutation columns in 8 sets.  k = 7: bytes == oracle prover;
    k = 14: accepted by the oracle verifier, rejected with a wrong public input or a broken witness."""
    from oracle import cref
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, 16, lambda a: cref.to_mont(0, a))
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        eng = HP.ResidentEngine(params, pk.vk.domain)
        timings = {}
        HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(1), engine=eng)
        proof = HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(k), engine=eng, timings=timings)
        ovk = PR.VerifyingKey(ocs, o.EvaluationDomain(5, k), pk.vk.fixed_commitments, pk.vk.permutation_commitments,
                              pk.vk.transcript_repr)
        vparams = PR.ParamsVerifier(k, S_TOXIC)
        assert PR.verify_proof(vparams, ovk, [public], proof, pairing=(k == 7))
        assert not PR.verify_proof(vparams, ovk, [[public[0] + 1] + public[1:]], proof)
        # the host-API engine (numpy multiplicities, batch bound scan on the host) gives the same bytes
        assert HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(k),
                               engine=HP.Engine(params, pk.vk.domain)) == proof
        if k == 7:
            oparams = PR.Params(k, S_TOXIC)
            opk = PR.keygen(oparams, ocs, [dec(c) for c in fixed], [[(int(c), int(r)) for c, r in col] for col in mapping],
                            transcript_repr=pk.vk.transcript_repr)
            assert PR.create_proof(oparams, opk, [dec(c) for c in advice], [public], HP.SeededRng(k)) == proof
        bad = advice.copy()
        bad[50, 9] = enc([1])[0]                              # a lookup input outside the table
        with pytest.raises(HP.B2Error):
            HP.create_proof(params, pk, bad, [public], HP.SeededRng(k), engine=eng)
        bad = advice.copy()
        bad[2, 9] = enc([12345])[0]                           # a product cell
        assert not PR.verify_proof(vparams, ovk, [public], HP.create_proof(params, pk, bad, [public], HP.SeededRng(k),
                                                                           engine=eng))
        eng.free()
        print(f"zkwasm shape k={k} create_proof {sum(timings.values()):.4f} s: "
              + ", ".join(f"{a} {b:.4f}" for a, b in timings.items()))
    finally:
        params.free()


def test_max_bits_and_device_multiplicity_primitives(gpu):
    """b2_fr_max_bits_dev (find_max_scalar_bits) and b2_logup_multiplicity_dev (radix sort of the table + the
    reference's binary search) against the oracle's logup_multiplicity"""
    import ctypes
    import random
    from halo2_gpu_specific_b200.evaluation import DeviceBuffer
    rng = random.Random(4)
    n = 1 << 12
    for bits in (0, 1, 16, 33, 64, 65, 200, 254):
        vals = [rng.randrange(1 << bits) if bits else 0 for _ in range(n)]
        if bits:
            vals[rng.randrange(n)] = (1 << bits) - 1 if bits < 254 else R - 1
        buf = DeviceBuffer(n).upload(enc(vals))
        got = ctypes.c_uint32()
        gpu.check(gpu.lib().b2_fr_max_bits_dev(ctypes.c_void_p(buf.ptr), n, ctypes.byref(got)))
        assert got.value == max(v.bit_length() for v in vals)
        buf.free()
    usable = n - 6

    def check_case(table, inputs, expect_miss=False):
        mont = np.stack([enc(col) for col in inputs + [table]])
        comp = DeviceBuffer((len(inputs) + 1) * n).upload(mont)
        m = DeviceBuffer(n)
        try:
            if expect_miss:
                with pytest.raises(gpu.B2Error):
                    HP.logup_multiplicity_device(comp.ptr, len(inputs), comp.ptr + len(inputs) * n * 32, usable, n, m.ptr)
                return
            want = PR.logup_multiplicity([inputs], table, usable, n)
            largest = HP.logup_multiplicity_device(comp.ptr, len(inputs), comp.ptr + len(inputs) * n * 32, usable, n, m.ptr)
            assert dec(m.download()) == want and largest == max(want)
        finally:
            comp.free(); m.free()

    # a table that repeats a few full-width values many times: which of the equal rows takes the count is decided by
    # the probe sequence of binary_search_by_key
    pool = [rng.randrange(R) for _ in range(40)] + [0, 1, 2]
    table = [rng.choice(pool) for _ in range(n)]
    check_case(table, [[rng.choice(table[:usable]) for _ in range(n)] for _ in range(3)])
    # distinct full-width values (theta-compressed multi-column tables): one limb decides the order
    table = [rng.randrange(R) for _ in range(n)]
    check_case(table, [[rng.choice(table[:usable]) for _ in range(n)] for _ in range(2)])
    # a 16-bit range table with zero padding (two radix passes, many equal rows)
    table = [i if i < 3000 else 0 for i in range(n)]
    check_case(table, [[rng.randrange(3000) if rng.random() < 0.8 else 0 for _ in range(n)]])
    # different keys that agree on the most significant differing limb: the lower limbs must be sorted too
    hi = [rng.randrange(1 << 60) << 192 for _ in range(8)]
    table = [rng.choice(hi) + rng.randrange(1 << 130) for _ in range(n)]
    check_case(table, [[rng.choice(table[:usable]) for _ in range(n)]])
    # all rows equal; no inputs at all
    check_case([7] * n, [[7] * n])
    check_case(table, [])
    # an input value that is not in the table (only the blinding rows hold it)
    table = list(range(n))
    bad = [5] * n
    bad[17] = usable + 1
    check_case(table, [bad], expect_miss=True)


@pytest.mark.parametrize("engine_kind", ["resident", "host_api"])
def test_shplonk_proof_bytes_match_oracle(gpu, engine_kind):
    """create_proof_with_shplonk on the engine == oracle bytes; accepted by the oracle's SHPLONK verifier (pairing)"""
    k = 5
    fx = fxm.build(k=k, seed=11)
    ocs = fx["cs"]
    oparams = PR.Params(k, S_TOXIC)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)
    params, pk = engine_side(k, oparams, cs, np.stack([enc(c) for c in fx["fixed"]]),
                             np.array(fx["mapping"], dtype=np.int64), opk.vk.transcript_repr)
    try:
        inst = [fx["instance"][0][:4]]
        want = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(9), use_gwc=False)
        adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
        eng = (HP.ResidentEngine if engine_kind == "resident" else HP.Engine)(params, pk.vk.domain)
        got = HP.create_proof_with_shplonk(params, pk, adv, inst, HP.SeededRng(9), engine=eng)
        eng.free()
        assert got == want
        assert PR.verify_proof(oparams, opk.vk, inst, got, use_gwc=False, pairing=True)
    finally:
        params.free()


def test_shplonk_zkwasm_shape_k14(gpu):
    from oracle import cref
    k = 14
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, 16, lambda a: cref.to_mont(0, a))
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        eng = HP.ResidentEngine(params, pk.vk.domain)
        timings = {}
        HP.create_proof_with_shplonk(params, pk, advice.copy(), [public], HP.SeededRng(1), engine=eng)
        proof = HP.create_proof_with_shplonk(params, pk, advice.copy(), [public], HP.SeededRng(2), engine=eng, timings=timings)
        eng.free()
        ovk = PR.VerifyingKey(ocs, o.EvaluationDomain(5, k), pk.vk.fixed_commitments, pk.vk.permutation_commitments,
                              pk.vk.transcript_repr)
        vparams = PR.ParamsVerifier(k, S_TOXIC)
        assert PR.verify_proof(vparams, ovk, [public], proof, use_gwc=False)
        assert not PR.verify_proof(vparams, ovk, [[public[0] + 1] + public[1:]], proof, use_gwc=False)
        print(f"shplonk zkwasm shape k={k} create_proof {sum(timings.values()):.4f} s, multiopen {timings['multiopen']:.4f} s")
    finally:
        params.free()


def test_sharded_engine_on_one_rank(gpu):
    """prover_sharded.ShardedResidentEngine without a process group (world = 1): same bytes as the plain engine; runs
    the on-device bound scan (B2_MAX_BITS_AUTO with resident columns), sub-blocks and the two-range transform path"""
    from halo2_gpu_specific_b200.prover_sharded import ShardedResidentEngine
    k = 6
    fx = fxm.build(k=k, seed=17)
    ocs = fx["cs"]
    oparams = PR.Params(k, S_TOXIC)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)
    params, pk = engine_side(k, oparams, cs, np.stack([enc(c) for c in fx["fixed"]]),
                             np.array(fx["mapping"], dtype=np.int64), opk.vk.transcript_repr)
    try:
        inst = [fx["instance"][0][:4]]
        adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
        want = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(5))
        eng = ShardedResidentEngine(params, pk.vk.domain)
        got = HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(5), engine=eng)
        # a forced split (as rank 1 of 3 would see it): this rank commits and transforms exactly its own column range
        # [lo, hi); the other columns would arrive in coefficient form from their owners (exchange_columns, a no-op
        # without a process group), so here they must still hold their Lagrange values
        eng._share = lambda count: (count // 3, count - count // 3)
        eng._gather = lambda local, count: local
        cols = fx["perm_z"] + [fx["shuffle_z"][0]]
        lo, hi = eng._share(len(cols))
        assert 0 < lo < hi < len(cols)
        z = eng.put(np.ascontiguousarray(np.stack([enc(p) for p in cols])))
        pts = eng.commit_lagrange_and_ifft(z)
        assert len(pts) == hi - lo
        assert pts == [oparams.commit_lagrange(c) for c in cols[lo:hi]]
        d = fx["domain"]
        want_rows = [enc(d.lagrange_to_coeff(c)) if lo <= i < hi else enc(c) for i, c in enumerate(cols)]
        assert np.array_equal(eng.get(z), np.stack(want_rows))
        eng.free()
        assert got == want
    finally:
        params.free()


@pytest.mark.parametrize("seed", range(16))
def test_random_circuits_proof_bytes_match_oracle(gpu, seed):
    """tests/test_prover_fuzz.py's generator (random gate expressions with constants, negation, scaling and rotations,
    a permutation over advice / fixed / instance, a lookup and a shuffle, satisfiable by construction) through BOTH
    device engines: the proof bytes must equal the oracle prover's; GWC and SHPLONK alternate."""
    import test_prover_fuzz as F
    oparams = PR.Params(F.K, F.S_TOXIC)
    params = h2.Params(F.K, oparams.g, oparams.g_lagrange)
    try:
        cs, fixed, advice, instance, mapping = F.build(seed)
        opk = PR.keygen(oparams, cs, fixed, mapping)
        inst = [instance[0][:3]]
        gwc = seed % 2 == 0
        want = PR.create_proof(oparams, opk, advice, inst, HP.SeededRng(seed), use_gwc=gwc)
        pk = HP.keygen(params, HP.ConstraintSystem.like(cs), np.stack([enc(c) for c in fixed]),
                       np.array(mapping, dtype=np.int64), transcript_repr=opk.vk.transcript_repr)
        adv = np.ascontiguousarray(np.stack([enc(c) for c in advice]))
        for kind in (HP.ResidentEngine, HP.Engine):
            eng = kind(params, pk.vk.domain)
            try:
                got = HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(seed), engine=eng, use_gwc=gwc)
            finally:
                eng.free()
            assert got == want, f"seed {seed}, {kind.__name__}"
    finally:
        params.free()


@pytest.mark.parametrize("engine_kind", ["resident", "host_api"])
@pytest.mark.parametrize("use_gwc", [True, False])
def test_multi_circuit_proof_bytes_match_oracle(gpu, engine_kind, use_gwc):
    """create_proof_ext(circuits: &[C], instances: &[&[&[Fr]]]) (plonk/prover.rs:206-222) on the device: two instances
    of the fixture circuit in one proof; bytes equal to the oracle's, which folds both through one evaluate_h
    accumulator as the reference does, and accepted by the oracle's multi-instance verifier"""
    k, seed = 5, 11
    fx, fx2 = fxm.build(k=k, seed=seed), fxm.build(k=k, seed=seed + 100)
    ocs = fx["cs"]
    oparams = PR.Params(k, S_TOXIC)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)
    params, pk = engine_side(k, oparams, cs, np.stack([enc(c) for c in fx["fixed"]]),
                             np.array(fx["mapping"], dtype=np.int64), opk.vk.transcript_repr)
    try:
        advs = [fx["advice"], fx2["advice"]]
        insts = [[fx["instance"][0][:4]], [fx2["instance"][0][:4]]]
        want = PR.create_proof_multi(oparams, opk, advs, insts, HP.SeededRng(9), use_gwc=use_gwc)
        dev_advs = [np.ascontiguousarray(np.stack([enc(c) for c in a])) for a in advs]
        eng = (HP.ResidentEngine if engine_kind == "resident" else HP.Engine)(params, pk.vk.domain)
        try:
            got = HP.create_proof_multi(params, pk, dev_advs, insts, HP.SeededRng(9), engine=eng, use_gwc=use_gwc)
        finally:
            eng.free()
        assert got == want
        assert PR.verify_proof_multi(oparams, opk.vk, insts, got, use_gwc=use_gwc, pairing=use_gwc)
        assert not PR.verify_proof_multi(oparams, opk.vk, insts[::-1], got, use_gwc=use_gwc)
    finally:
        params.free()


def test_early_advice_transforms_do_not_change_the_proof(gpu):
    """ResidentEngine turns the advice columns into coefficient form and coset evaluations on a side stream while the
    later columns are still being uploaded (EARLY_TRANSFORMS); the proof bytes are those of the plain schedule, twice
    in a row on the same engine (buffers recycled through the pool)"""
    from oracle import cref
    k = 9
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, 4, lambda a: cref.to_mont(0, a))
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        plain = HP.ResidentEngine(params, pk.vk.domain)
        plain.EARLY_TRANSFORMS = False
        want = HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(3), engine=plain)
        assert not plain._early
        plain.free()
        eng = HP.ResidentEngine(params, pk.vk.domain)
        seen = []
        orig = eng._early_transforms
        eng._early_transforms = lambda *a: (seen.append(a[3:]), orig(*a))[1]
        for _ in range(2):
            assert HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(3), engine=eng) == want
        assert len(seen) == 2 * 8 and seen[0] == (0, 8)          # 64 advice columns in 8 groups, per proof
        assert HP.create_proof_with_shplonk(params, pk, advice.copy(), [public], HP.SeededRng(4), engine=eng) == \
            HP.create_proof_with_shplonk(params, pk, advice.copy(), [public], HP.SeededRng(4), engine=HP.Engine(params, pk.vk.domain))
        eng.free()
    finally:
        params.free()


def test_early_advice_transforms_with_two_circuit_instances(gpu):
    """create_proof_multi over two witnesses of the zkWasm-shaped circuit: each instance's 64 advice columns take the
    early-transform path (two entries in the engine's table, two sets of coset evaluations alive at once); bytes as
    without it, and the host-API engine agrees"""
    import zkwasm_shape_circuit as zk
    from oracle import cref
    k = 8
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, 4, lambda a: cref.to_mont(0, a))
    _, advice2, public2, _ = zk.build(k, lambda a: cref.to_mont(0, a), seed=k + 100)
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        advs = lambda: [advice.copy(), advice2.copy()]               # noqa: E731
        insts = [[public], [public2]]
        plain = HP.ResidentEngine(params, pk.vk.domain)
        plain.EARLY_TRANSFORMS = False
        want = HP.create_proof_multi(params, pk, advs(), insts, HP.SeededRng(11), engine=plain)
        plain.free()
        eng = HP.ResidentEngine(params, pk.vk.domain)
        got = HP.create_proof_multi(params, pk, advs(), insts, HP.SeededRng(11), engine=eng)
        assert len(eng._early) == 0            # released with the proof's buffers
        eng.free()
        assert got == want
        host = HP.Engine(params, pk.vk.domain)
        assert HP.create_proof_multi(params, pk, advs(), insts, HP.SeededRng(11), engine=host) == want
    finally:
        params.free()


def test_point_range_commit_pieces_on_one_gpu(gpu):
    """the device halves of prover_sharded.ShardedCommits.commit_by_point_range (SURVEY 8(e) option (i)) without a
    process group: the partial MSMs of every column over the three point ranges a 3-rank run would use
    (msm_partials: b2_msm_dev at an offset into the window table), stacked rank-major as the all-gather leaves them,
    summed in one launch (sum_partials: b2_g1_sum_groups_dev) -- equal to the whole-column commitments, for both bases,
    with an empty range, a zero column and a column that leaves one rank an identity partial"""
    import torch
    from halo2_gpu_specific_b200 import parallel
    from halo2_gpu_specific_b200.prover_sharded import ShardedResidentEngineQ
    from oracle import cref
    k = 11
    n = 1 << k
    oparams = PR.Params(k, S_TOXIC)
    params = h2.Params(k, oparams.g, oparams.g_lagrange)
    try:
        eng = ShardedResidentEngineQ(params, h2.EvaluationDomain(5, k))
        cols = np.stack([cref.random_fr_mont(n, 0x7A0 + i) for i in range(3)] + [np.zeros((n, 4), dtype=np.uint64)] * 2)
        cols[4, 5] = enc([7])[0]                       # column 4: two lone scalars that fall into different ranks' ranges,
        cols[4, n - 3] = enc([o.R_MOD - 1])[0]         # so one rank's partial of it is the identity
        block = eng.put(np.ascontiguousarray(cols))
        for basis, whole in (("g", eng.commit(block)), ("g_lagrange", eng.commit_lagrange(block))):
            for world in (3, 1):
                parts = [eng.msm_partials(basis, block, *parallel.shard_range(n, world, r), 254) for r in range(world)]
                parts.append(eng.msm_partials(basis, block, n, n, 254))       # a rank beyond the data: identities
                gathered = torch.cat(parts).contiguous()
                assert eng.sum_partials(gathered, world + 1, block.count) == whole
            assert whole[3] is None and whole[4] is not None
        want = [oparams.commit(c) for c in [o.fr_decode(cols[0]), o.fr_decode(cols[4])]]
        assert [eng.commit(block)[0], eng.commit(block)[4]] == want
        eng.free()
    finally:
        params.free()


def test_early_transforms_of_a_column_share(gpu):
    """what a rank of the sharded prover does with its share of the advice columns (put_share_and_commit: groups of
    columns committed on the way in, coefficient forms and coset evaluations of THOSE columns made behind the upload),
    on one GPU: columns [16, 40) go that way, the others arrive plainly (as a peer's would); lagrange_to_coeff and
    evaluate_h_blocks must combine the early results with what they compute themselves -- same proof bytes"""
    from oracle import cref
    from halo2_gpu_specific_b200.prover_sharded import ShardedResidentEngineQ
    k = 8
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, 4, lambda a: cref.to_mont(0, a))
    params = h2.Params.unsafe_setup(k, S_TOXIC)
    try:
        pk = HP.keygen(params, cs, fixed, mapping)
        plain = HP.ResidentEngine(params, pk.vk.domain)
        plain.EARLY_TRANSFORMS = False
        want = HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(21), engine=plain)
        plain.free()
        eng = ShardedResidentEngineQ(params, pk.vk.domain)          # no process group: world = 1
        lo, hi = 16, 40

        def put(host, max_bits):
            bits = 0xFFFFFFFF if max_bits is None else max_bits
            block = eng.alloc(host.shape[0])
            mid = eng.put_share_and_commit(block, host, lo, hi, max_bits)
            assert block.ptr in eng._early and (eng._early[block.ptr]["lo"], eng._early[block.ptr]["hi"]) == (lo, hi)
            first = eng._commit(params.g_lagrange, host[:lo].ctypes.data, eng.sub_block(block, 0, lo), bits, False)
            last = eng._commit(params.g_lagrange, host[hi:].ctypes.data, eng.sub_block(block, hi, host.shape[0]), bits, False)
            return block, first + mid + last

        eng.put_and_commit_lagrange = put
        for _ in range(2):
            assert HP.create_proof(params, pk, advice.copy(), [public], HP.SeededRng(21), engine=eng) == want
        eng.free()
    finally:
        params.free()
