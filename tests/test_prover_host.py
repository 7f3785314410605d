"""CPU tests of the prover mirror's HOST logic (halo2_gpu_specific_b200/plonk.py, transcript.py): the same
create_proof / keygen code that drives the device engine is run here over a test double that answers every numeric
call with the CPU oracle (tests/oracle_engine.py), and must produce the oracle prover's bytes (oracle/prover.py,
written independently from the reference) -- transcript order, RNG order, query grouping, multiplicities and the
Evaluator graph are all host decisions.  The device engine itself is covered by tests/test_gpu_prover.py."""
import os
import random
import sys

import numpy as np
import pytest

import plonk_fixture as fxm
from oracle import bn254 as o
from oracle import plonk as P
from oracle import prover as PR
from oracle_engine import OracleEngine

from halo2_gpu_specific_b200 import plonk as HP
from halo2_gpu_specific_b200 import transcript as HT

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import plonk_bench_circuit as bench_circuit  # noqa: E402

R = o.R_MOD
enc, dec = o.fr_encode, o.fr_decode
S_TOXIC = 0x2B200B200B200B200B200B200B200B2001


class HostParams:
    """what keygen / create_proof read from Params when the engine is injected"""

    def __init__(self, k):
        self.k, self.n = k, 1 << k


def both_sides(k, seed):
    fx = fxm.build(k=k, seed=seed)
    ocs = fx["cs"]
    oparams = PR.Params(k, S_TOXIC)
    opk = PR.keygen(oparams, ocs, fx["fixed"], fx["mapping"])
    cs = HP.ConstraintSystem.like(ocs)
    eng = OracleEngine(oparams, opk.vk.domain, ocs)
    pk = HP.keygen(HostParams(k), cs, np.stack([enc(c) for c in fx["fixed"]]), np.array(fx["mapping"], dtype=np.int64),
                   engine=eng, transcript_repr=opk.vk.transcript_repr)
    return fx, oparams, opk, cs, eng, pk


def test_evaluator_graph_matches_reference_construction():
    """Evaluator::new: same constants, rotations, calculations (order and sharing), value parts, lookup and
    shuffle results as the oracle's restatement of evaluation.rs:307-448"""
    for seed in (3, 11):
        fx = fxm.build(k=5, seed=seed)
        parts = HP.evaluator_parts(HP.ConstraintSystem.like(fx["cs"]))
        ev = P.Evaluator.new(fx["cs"])
        assert parts["rotations"] == ev.rotations and parts["constants"] == ev.constants
        assert parts["calculations"] == ev.calculations
        assert parts["value_parts"] == ev.value_parts
        assert parts["lookup_results"] == ev.lookup_results
        assert parts["shuffle_results"] == ev.shuffle_results


def test_evaluator_graph_simplifications():
    cs = HP.ConstraintSystem(1, 2, 0, degree=3)
    a, b = ("Advice", 0, 0), ("Advice", 1, 0)
    cs.gates.append([("Product", a, b), ("Product", b, a), ("Sum", ("Product", a, b), ("Constant", 0)),
                     ("Scaled", a, 1), ("Scaled", b, 0), ("Negated", ("Constant", 1)),
                     ("Sum", ("Constant", 0), ("Negated", b))])
    p = HP.evaluator_parts(cs)
    assert p["constants"][:2] == [0, 1] and p["constants"][2] == R - 1
    assert sum(1 for c in p["calculations"] if c[0] == "Mul") == 1          # a*b and b*a share one slot
    vp = p["value_parts"]
    assert vp[0] == vp[1] == vp[2]
    assert vp[4] == ("Constant", 0) and vp[5] == ("Constant", 2)
    assert vp[6] == vp[3 + 0] or p["calculations"][vp[6][1]] == ("Store", ("Advice", 1, 0))   # (sic) 0 - b -> b


def test_queries_follow_the_documented_rule():
    fx = fxm.build(k=5, seed=3)
    cs = HP.ConstraintSystem.like(fx["cs"])
    q = cs.queries()
    assert q == PR.collect_queries(fx["cs"])
    assert q["Advice"][:3] == [(0, 0), (1, 0), (2, 0)] and (0, 1) in q["Advice"] and (2, -1) in q["Advice"]
    assert q["Fixed"][0] == (5, 0) and q["Instance"] == [(0, 0)]
    explicit = HP.ConstraintSystem(**bench_circuit.constraint_system_args())
    assert explicit.queries()["Fixed"] == [(1, 0), (2, 0), (3, 0), (0, 0)]


def test_keygen_matches_oracle():
    fx, oparams, opk, cs, eng, pk = both_sides(5, 11)
    assert pk.vk.fixed_commitments == opk.vk.fixed_commitments
    assert pk.vk.permutation_commitments == opk.vk.permutation_commitments
    for got, want in ((pk.sigmas, opk.sigmas), (pk.sigma_polys, opk.sigma_polys), (pk.fixed_polys, opk.fixed_polys)):
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert np.array_equal(a, enc(b))
    assert np.array_equal(pk.l0, enc(opk.l0))
    assert np.array_equal(pk.l_last, enc(opk.l_last))
    assert np.array_equal(pk.l_active_row, enc(opk.l_active_row))
    # without an explicit scalar the key hashes its own description, deterministically
    a = HP.VerifyingKey(cs, pk.vk.domain, pk.vk.fixed_commitments, pk.vk.permutation_commitments, None)
    b = HP.VerifyingKey(cs, pk.vk.domain, pk.vk.fixed_commitments, pk.vk.permutation_commitments, None)
    assert a.transcript_repr == b.transcript_repr != 0


@pytest.mark.parametrize("k,seed,rng_seed", [(5, 11, 1), (5, 3, 7), (6, 17, 2)])
def test_create_proof_bytes_match_oracle_and_verify(k, seed, rng_seed):
    fx, oparams, opk, cs, eng, pk = both_sides(k, seed)
    inst = [fx["instance"][0][:4]]
    want = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(rng_seed))
    adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
    timings = {}
    got = HP.create_proof(HostParams(k), pk, adv, inst, HP.SeededRng(rng_seed), engine=eng, timings=timings)
    assert got == want
    assert PR.verify_proof(oparams, opk.vk, inst, got)
    assert set(timings) == {"instance", "advice", "lookup_m", "z_columns", "vanishing_commit", "h_poly", "evaluations",
                            "multiopen"}


def test_create_proof_argument_checks():
    fx, oparams, opk, cs, eng, pk = both_sides(5, 11)
    adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
    with pytest.raises(HP.B2Error):
        HP.create_proof(HostParams(5), pk, adv, [], HP.SeededRng(1), engine=eng)                  # InvalidInstances
    with pytest.raises(HP.B2Error):
        HP.create_proof(HostParams(5), pk, adv, [[1] * 27], HP.SeededRng(1), engine=eng)          # InstanceTooLarge
    with pytest.raises(HP.B2Error):
        HP.create_proof(HostParams(5), pk, adv[:-1], [[1]], HP.SeededRng(1), engine=eng)
    bad = adv.copy()
    bad[7, 3] = enc([1])[0]                                                                      # not in the table
    with pytest.raises(HP.B2Error):
        HP.create_proof(HostParams(5), pk, bad, [fx["instance"][0][:4]], HP.SeededRng(1), engine=eng)


def test_benches_plonk_circuit_proves_and_verifies():
    """BASELINE config 4's circuit (benches/plonk.rs) at k = 6: no lookups, shuffles or instances; explicit query
    order as the reference's configure() produces it"""
    k = 6
    cs = HP.ConstraintSystem(**bench_circuit.constraint_system_args())
    fixed, advice, mapping = bench_circuit.build(k)
    ocs = P.ConstraintSystem(4, 3, 0, degree=5, blinding_factors=5)
    ocs.gates, ocs.permutation_columns = cs.gates, cs.permutation_columns
    ocs.advice_queries, ocs.fixed_queries, ocs.instance_queries = cs.advice_queries, cs.fixed_queries, cs.instance_queries
    oparams = PR.Params(k, S_TOXIC)
    omap = [[(int(c), int(r)) for c, r in col] for col in mapping]
    opk = PR.keygen(oparams, ocs, [dec(c) for c in fixed], omap)
    eng = OracleEngine(oparams, opk.vk.domain, ocs)
    pk = HP.keygen(HostParams(k), cs, fixed, mapping, engine=eng, transcript_repr=opk.vk.transcript_repr)
    want = PR.create_proof(oparams, opk, [dec(c) for c in advice], [], HP.SeededRng(5))
    got = HP.create_proof(HostParams(k), pk, advice.copy(), [], HP.SeededRng(5), engine=eng, advice_max_bits=254)
    assert got == want
    assert PR.verify_proof(oparams, opk.vk, [], got)
    # 3 advice + 1 z... : A=3, P=1 sets (3 columns, chunk 3), vanishing 1 + 4 h pieces; evals 3+4+1+3+2; W for rot 0, 1
    assert len(got) == 32 * ((3 + 1 + 1 + 4) + (3 + 4 + 1 + 3 + 2) + 2)
    # an unsatisfied witness must not verify
    broken = advice.copy()
    broken[2, 4] = enc([5])[0]
    bad = HP.create_proof(HostParams(k), pk, broken, [], HP.SeededRng(5), engine=eng)
    assert not PR.verify_proof(oparams, opk.vk, [], bad)


def test_logup_multiplicity_matches_scalar_restatement():
    """numpy sort / searchsorted / probe simulation vs the scalar restatement of binary_search_by_key, on tables
    with long runs of repeated values"""
    rng = random.Random(5)
    for trial in range(20):
        n = 64
        usable = n - 6
        distinct = [rng.randrange(R) for _ in range(rng.randrange(1, 12))] + [0, 1, R - 1]
        table = [rng.choice(distinct) for _ in range(n)]
        inputs = [[rng.choice(table[:usable]) for _ in range(n)] for _ in range(3)]
        want = PR.logup_multiplicity([inputs[:2], inputs[2:]], table, usable, n)
        canon = lambda col: np.array([o._to_limbs(v) for v in col], dtype=np.uint64)      # noqa: E731
        got = HP.logup_multiplicity([canon(c) for c in inputs], canon(table), usable, n)
        assert got.tolist() == want
    with pytest.raises(HP.B2Error):
        HP.logup_multiplicity([canon([2] * n)], canon([3] * n), usable, n)


def test_transcript_matches_oracle_transcript():
    rng = random.Random(9)
    a, b = HT.Blake2bWrite(), PR.Blake2bWrite()
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(R)) for _ in range(4)]
    for step in range(40):
        op = rng.randrange(5)
        if op == 0:
            assert a.squeeze_challenge() == b.squeeze_challenge()
        elif op == 1:
            p = rng.choice(pts)
            a.write_point(o.g1_jacobian_encode(p))           # as the engine returns it: Jacobian, Z = 1
            b.write_point(p)
        elif op == 2:
            p = rng.choice(pts)
            a.common_point(o.g1_affine_encode([p])[0])
            b.common_point(p)
        elif op == 3:
            s = rng.randrange(R)
            a.write_scalar(enc([s])[0])                       # Montgomery limbs
            b.write_scalar(s)
        else:
            s = rng.randrange(R)
            a.common_scalar(s)
            b.common_scalar(s)
    assert a.finalize() == b.finalize() and a.squeeze_challenge() == b.squeeze_challenge()
    with pytest.raises(HT.B2Error):
        a.write_point(np.zeros(12, dtype=np.uint64))           # identity: "cannot write points at infinity"
    assert HT.point_from_engine(np.zeros(8, dtype=np.uint64)) is None
    assert HT.g1_to_bytes(None) == bytes(32)
    assert HT.g1_to_bytes((1, 2)) == o.g1_to_bytes((1, 2)) and HT.g1_to_bytes((5, 3), 6) == o.g1_to_bytes((5, 3), 6)


def test_seeded_rng_is_reproducible_and_in_range():
    a, b = HP.SeededRng(4), HP.SeededRng(4)
    assert np.array_equal(a.fr_vec(100), b.fr_vec(100)) and np.array_equal(a.u16_vec(9), b.u16_vec(9))
    assert all(o._limbs_to_int(r) < R for r in a.fr_vec(200))
    assert int(a.u16_vec(1000).max()) < 1 << 16
    x = HP.OsRng()
    assert x.fr_vec(3).shape == (3, 4) and x.u16_vec(5).shape == (5,) and x.u64_vec(2).dtype == np.uint64


def test_device_engine_refuses_to_run_without_a_gpu():
    """no CPU fallback: constructing the device engine (what create_proof does by default) needs CUDA"""
    from halo2_gpu_specific_b200 import _lib
    if _lib.lib().b2_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(HP.B2Error):
        HP.Engine(HostParams(5), None)
    fx, oparams, opk, cs, eng, pk = both_sides(5, 11)
    adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
    with pytest.raises(HP.B2Error):
        HP.create_proof(HostParams(5), pk, adv, [fx["instance"][0][:4]], HP.SeededRng(1))


def test_device_multiplicity_algorithm_matches_scalar_restatement():
    """csrc/lookup.cuh restated in Python (stable sort of the table's usable rows, then the pinned toolchain's
    binary_search_by loop per input value, the first Equal probe takes the count) against the oracle's
    logup_multiplicity, which simulates the probes on runs of equal keys: wide keys, narrow keys, long runs of repeated
    table values.  The kernels themselves are checked on the GPU (tests/test_gpu_prover.py)."""
    rng = random.Random(6)

    def kernel_algorithm(inputs, table, usable, n):
        order = sorted(range(usable), key=lambda i: table[i])            # stable
        skeys = [table[i] for i in order]
        m = [0] * n
        for col in inputs:
            for v in col[:usable]:
                size, left, right = usable, 0, usable
                hit = None
                while left < right:
                    mid = left + size // 2
                    if skeys[mid] == v:
                        hit = mid
                        break
                    if skeys[mid] < v:
                        left = mid + 1
                    else:
                        right = mid
                    size = right - left
                assert hit is not None
                m[order[hit]] += 1
        return m

    for trial in range(24):
        n = 128
        usable = n - 6
        wide = trial % 2 == 0
        pool = [rng.randrange(R) if wide else rng.randrange(1 << 40) for _ in range(rng.randrange(1, 20))] + [0, 1]
        table = [rng.choice(pool) for _ in range(n)]
        inputs = [[rng.choice(table[:usable]) for _ in range(n)] for _ in range(1 + trial % 3)]
        assert kernel_algorithm(inputs, table, usable, n) == PR.logup_multiplicity([inputs], table, usable, n)


def _zk_shape(k, extra_gates):
    import zkwasm_shape_circuit as zk
    from oracle import cref
    args = zk.constraint_system_args(extra_gates=extra_gates)
    cs = HP.ConstraintSystem(**args)
    ocs = P.ConstraintSystem(args["num_fixed"], args["num_advice"], args["num_instance"], degree=5, blinding_factors=5)
    ocs.gates, ocs.lookups, ocs.shuffles = cs.gates, cs.lookups, cs.shuffles
    ocs.permutation_columns = cs.permutation_columns
    fixed, advice, public, mapping = zk.build(k, lambda a: cref.to_mont(0, a), seed=k)
    return cs, ocs, fixed, advice, public, mapping


def test_zkwasm_shaped_circuit_proves_and_verifies():
    """BASELINE config 5's shape (64 advice, 32 fixed, 8 lookups / 12 input sets, 4 shuffles, 24 permutation columns)
    with a real witness at k = 6: host logic over the oracle-backed engine == oracle prover, and the verifier accepts"""
    k = 6
    cs, ocs, fixed, advice, public, mapping = _zk_shape(k, extra_gates=8)
    oparams = PR.Params(k, S_TOXIC)
    omap = [[(int(c), int(r)) for c, r in col] for col in mapping]
    opk = PR.keygen(oparams, ocs, [dec(c) for c in fixed], omap)
    eng = OracleEngine(oparams, opk.vk.domain, ocs)
    pk = HP.keygen(HostParams(k), cs, fixed, mapping, engine=eng, transcript_repr=opk.vk.transcript_repr)
    want = PR.create_proof(oparams, opk, [dec(c) for c in advice], [public], HP.SeededRng(3))
    got = HP.create_proof(HostParams(k), pk, advice.copy(), [public], HP.SeededRng(3), engine=eng)
    assert got == want
    assert PR.verify_proof(oparams, opk.vk, [public], got)
    assert not PR.verify_proof(oparams, opk.vk, [[public[0] + 1] + public[1:]], got)
    bad = advice.copy()
    bad[2, 9] = enc([12345])[0]                       # a product cell
    assert not PR.verify_proof(oparams, opk.vk, [public],
                               HP.create_proof(HostParams(k), pk, bad, [public], HP.SeededRng(3), engine=eng))


@pytest.mark.parametrize("k,seed", [(5, 11), (6, 17)])
def test_shplonk_proof_bytes_match_oracle_and_verify(k, seed):
    """create_proof_with_shplonk (plonk/prover.rs:1737-1757, poly/multiopen/shplonk/prover.rs): the engine-side
    formulation (one y-fold per rotation set, low-degree parts as host scalars) gives the oracle's bytes, which
    restate the reference's per-commitment formulation; the oracle's SHPLONK verifier accepts"""
    fx, oparams, opk, cs, eng, pk = both_sides(k, seed)
    inst = [fx["instance"][0][:4]]
    want = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(4), use_gwc=False)
    adv = np.ascontiguousarray(np.stack([enc(c) for c in fx["advice"]]))
    got = HP.create_proof_with_shplonk(HostParams(k), pk, adv, inst, HP.SeededRng(4), engine=eng)
    assert got == want
    assert PR.verify_proof(oparams, opk.vk, inst, got, use_gwc=False)
    assert PR.verify_proof(oparams, opk.vk, inst, got, use_gwc=False, pairing=(k == 5))
    assert not PR.verify_proof(oparams, opk.vk, [[inst[0][0] + 1] + inst[0][1:]], got, use_gwc=False)
    gwc = PR.create_proof(oparams, opk, fx["advice"], inst, HP.SeededRng(4))
    # same transcript up to the multiopen argument; 2 points (h1, h2) instead of one W per rotation (4 here)
    assert len(got) == len(gwc) - 64 and got[:len(got) - 64] == gwc[:len(got) - 64]


def test_degenerate_shapes_prove_and_verify():
    """circuits that lack whole arguments: (a) gates only, no permutation / lookup / shuffle / instance;
    (b) one lookup and nothing else besides a gate"""
    k = 5
    n = 1 << k
    usable = n - 6
    rng = random.Random(8)
    oparams = PR.Params(k, S_TOXIC)
    A, F = ("Advice", 0, 0), ("Fixed", 0, 0)
    sq = ("Product", F, ("Sum", ("Product", A, A), ("Negated", ("Advice", 1, 0))))
    for with_lookup in (False, True):
        lookups = [{"table_expressions": [("Fixed", 1, 0)], "input_expressions_sets": [[[("Advice", 1, 0)]]]}] if with_lookup else []
        # cs.degree(): 3 for the gate alone, 4 with the lookup (active rows * table * input * z); a larger value would
        # leave an h(X) piece identically zero, which neither prover can commit to ("points at infinity")
        degree = 4 if with_lookup else 3
        cs = HP.ConstraintSystem(2, 2, 0, degree=degree, blinding_factors=5, gates=[[sq]], lookups=lookups)
        ocs = P.ConstraintSystem(2, 2, 0, degree=degree, blinding_factors=5)
        ocs.gates, ocs.lookups = cs.gates, cs.lookups
        a0 = [rng.randrange(R) for _ in range(n)]
        a1 = [v * v % R for v in a0]
        fixed = [[1 if r < usable else 0 for r in range(n)], [a1[(r * 7) % usable] for r in range(n)]]
        advice = [a0, a1]
        opk = PR.keygen(oparams, ocs, fixed, [])
        eng = OracleEngine(oparams, opk.vk.domain, ocs)
        pk = HP.keygen(HostParams(k), cs, np.stack([enc(c) for c in fixed]), np.zeros((0, n, 2), dtype=np.int64), engine=eng,
                       transcript_repr=opk.vk.transcript_repr)
        for gwc in (True, False):
            want = PR.create_proof(oparams, opk, advice, [], HP.SeededRng(2), use_gwc=gwc)
            got = HP.create_proof(HostParams(k), pk, np.stack([enc(c) for c in advice]), [], HP.SeededRng(2), engine=eng,
                                  use_gwc=gwc)
            assert got == want
            assert PR.verify_proof(oparams, opk.vk, [], got, use_gwc=gwc)
        bad = [list(a0), list(a1)]
        bad[1][3] = (bad[1][3] + 1) % R
        if not with_lookup:
            assert not PR.verify_proof(oparams, opk.vk, [], PR.create_proof(oparams, opk, bad, [], HP.SeededRng(2)))


def test_degree_and_blinding_factors_rules():
    """ConstraintSystem::degree / blinding_factors (circuit.rs:1861-1944) computed from the constraint system"""
    args = bench_circuit.constraint_system_args()
    args.update(degree=None, blinding_factors=None, minimum_degree=5)          # benches/plonk.rs: set_minimum_degree(5)
    cs = HP.ConstraintSystem(**args)
    assert cs.compute_degree() == 5 and cs.degree() == 5 and cs.blinding_factors() == 5
    args.update(minimum_degree=None)
    assert HP.ConstraintSystem(**args).degree() == 3                             # gate a*b*sm: degree 3
    fx = fxm.build(k=5, seed=3)
    like = HP.ConstraintSystem.like(fx["cs"])
    assert like.compute_degree() == 4                                            # logup: max(4, 2 + 1 + 1)
    assert like.compute_blinding_factors() == 5                                  # advice 0 and 2 are queried twice
    import zkwasm_shape_circuit as zk
    z = HP.ConstraintSystem(**zk.constraint_system_args(extra_gates=30))
    assert z.compute_degree() == 4 and z.degree() == 5                           # extras have degree 4; the shape fixes 5
    A = ("Advice", 0, 0)
    deg = HP.ConstraintSystem.expression_degree
    assert deg(("Product", ("Sum", A, ("Constant", 1)), ("Scaled", ("Negated", A), 3))) == 2 and deg(("Constant", 7)) == 0


def test_buffer_pool_recycles_and_trims(monkeypatch):
    """evaluation.BufferPool / DeviceBuffer: exact-size reuse, no allocation in the steady state, trim + retry when the
    device is out of memory with blocks parked in the pool (driver calls replaced by a fake allocator)"""
    import ctypes
    from halo2_gpu_specific_b200 import evaluation as E

    class FakeLib:
        def __init__(self):
            self.next, self.live, self.allocs, self.frees, self.capacity = 0x1000, {}, 0, 0, 10 * 32 * 100

        def b2_dev_alloc(self, nbytes, out):
            if sum(self.live.values()) + nbytes > self.capacity:
                return -3
            self.next += 0x10000
            self.live[self.next] = nbytes
            ctypes.cast(out, ctypes.POINTER(ctypes.c_void_p))[0] = self.next
            self.allocs += 1
            return 0

        def b2_dev_free(self, p):
            del self.live[p.value]
            self.frees += 1
            return 0

        def b2_last_error(self):
            return b"out of memory"

    fake = FakeLib()
    monkeypatch.setattr(E, "lib", lambda: fake)
    from halo2_gpu_specific_b200 import _lib
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    pool = E.BufferPool()
    prev = E.set_active_pool(pool)
    try:
        a, b = E.DeviceBuffer(100), E.DeviceBuffer(200)
        pa, pb = a.ptr, b.ptr
        a.free(); b.free()
        assert fake.frees == 0 and pool.cached_bytes == 300 * 32
        c = E.DeviceBuffer(200)                      # same size: recycled, no driver call
        assert c.ptr == pb and fake.allocs == 2
        d = E.DeviceBuffer(100)
        assert d.ptr == pa and pool.cached_bytes == 0
        c.free(); d.free()
        big = E.DeviceBuffer(800)                    # does not fit while 300 elements are parked: trim, retry
        assert fake.frees == 2 and pool.cached_bytes == 0 and fake.live == {big.ptr: 800 * 32}
        big.free()
        with pytest.raises(HP.B2Error):
            E.DeviceBuffer(2000)                     # genuinely too large
    finally:
        E.set_active_pool(prev)
    e = E.DeviceBuffer(10)                           # no pool: plain alloc / free
    n = fake.frees
    e.free()
    assert fake.frees == n + 1
    pool.trim()
    assert not fake.live


def test_small_host_helpers_match_oracle():
    """lagrange_interpolate (arithmetic.rs:848-906), the vanishing streams' generator and canonical_max_bits against
    the oracle's independent restatements"""
    rng = random.Random(12)
    for m in range(1, 6):
        pts = rng.sample(range(1, 1 << 60), m)
        evs = [rng.randrange(R) for _ in range(m)]
        got = HP.lagrange_interpolate(pts, evs)
        assert got == o.lagrange_interpolate(pts, evs)
        for p, e in zip(pts, evs):
            assert o.eval_polynomial(got, p) == e
    key = bytes(range(100, 132))
    a_lo, a_hi, u, b_lo, b_hi, v = HP.vanishing_streams(key, 16)
    val = lambda limbs: sum(int(x) << (64 * l) for l, x in enumerate(limbs))          # noqa: E731
    for i in (0, 1, 7, 15):
        blk = [PR.chacha20_block(key, 3 * i + j) for j in range(3)]
        assert val(a_lo[i]) + (val(a_hi[i]) << 256) == int.from_bytes(blk[0], "little")
        assert val(b_lo[i]) + (val(b_hi[i]) << 256) == int.from_bytes(blk[1], "little")
        assert int(u[i]) == int.from_bytes(blk[2][:8], "little") and int(v[i]) == int.from_bytes(blk[2][8:16], "little")
    canon = np.array([o._to_limbs(v) for v in (0, 1, 255, 1 << 64, (1 << 130) + 5)], dtype=np.uint64)
    assert HP.canonical_max_bits(canon) == 131 and HP.canonical_max_bits(canon[:3]) == 8
    assert HP.canonical_max_bits(canon[:1]) == 0 and HP.canonical_max_bits(np.zeros((0, 4), np.uint64)) == 0


def test_chacha20_three_implementations_agree():
    """The generator of the vanishing argument's random polynomial: RFC 8439's block-function vector (2.3.2), and on
    random keys / counters the oracle's scalar restatement, the package's vectorised one, the HOST BUILD OF THE DEVICE
    SOURCE (csrc/chacha.cuh through host/chacha_selftest.cpp) and an independent library (cryptography's ChaCha20)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "halo2_gpu_specific_b200", "host", "chacha_selftest")
    if not os.path.exists(exe):
        import __graft_entry__ as g
        g.build()
    rfc_key = bytes(range(32))
    want = bytes.fromhex("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                         "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    assert PR.chacha20_block(rfc_key, 1, (0x09000000, 0x4A000000, 0)) == want
    out = subprocess.run([exe, rfc_key.hex(), "1", "1", "0x09000000", "0x4a000000", "0"], capture_output=True, text=True,
                         check=True).stdout.split()
    assert bytes.fromhex(out[0]) == want
    rng = random.Random(20)
    for first, count in ((0, 7), (3 * ((1 << 22) - 2), 6), ((1 << 32) - 3, 3)):
        key = bytes(rng.randrange(256) for _ in range(32))
        scalar = [PR.chacha20_block(key, first + j) for j in range(count)]
        vec = HP.chacha20_blocks(key, first, count)
        assert [vec[j].astype("<u4").tobytes() for j in range(count)] == scalar
        dev_src = subprocess.run([exe, key.hex(), str(first), str(count)], capture_output=True, text=True,
                                 check=True).stdout.split()
        assert [bytes.fromhex(h) for h in dev_src] == scalar
        try:
            from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
        except ImportError:
            continue
        for j in range(count):
            nonce16 = ((first + j) & 0xFFFFFFFF).to_bytes(4, "little") + bytes(12)
            ks = Cipher(algorithms.ChaCha20(key, nonce16), mode=None).encryptor().update(bytes(64))
            assert ks == scalar[j]


def test_cpp_transcript_matches_python(tmp_path):
    """host/halo2_b200_transcript.hpp (Blake2b from RFC 7693, Blake2bWrite, Challenge255 / from_bytes_wide) against the
    oracle's and the package's Python transcripts on a random sequence of operations; host-only, no GPU needed"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "transcript_selftest")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "include"),
                           "-I", os.path.join(root, "halo2_gpu_specific_b200", "host"), "-o", exe,
                           os.path.join(root, "halo2_gpu_specific_b200", "host", "transcript_selftest.cpp")])
    rng = random.Random(21)
    ref, py = PR.Blake2bWrite(), HT.Blake2bWrite()
    pts = [o.g1_mul(o.G1_GEN, rng.randrange(R)) for _ in range(5)]
    lines, want = [], []
    for step in range(300):            # long enough to cross many 128-byte block boundaries
        op = rng.randrange(5)
        if op == 0:
            lines.append("C")
            c = ref.squeeze_challenge()
            assert py.squeeze_challenge() == c
            want.append("C " + enc([c])[0].tobytes().hex())
        elif op in (1, 2):
            s = rng.choice([0, 1, R - 1, rng.randrange(R)])
            lines.append(("S " if op == 1 else "s ") + enc([s])[0].tobytes().hex())
            (ref.common_scalar if op == 1 else ref.write_scalar)(s)
            (py.common_scalar if op == 1 else py.write_scalar)(s)
        else:
            p = rng.choice(pts)
            lines.append(("P " if op == 3 else "p ") + o.g1_affine_encode([p])[0].tobytes().hex())
            (ref.common_point if op == 3 else ref.write_point)(p)
            (py.common_point if op == 3 else py.write_point)(p)
    lines.append("P " + bytes(64).hex())                                   # identity: refused, state untouched
    want.append("E cannot write points at infinity to the transcript")
    lines.append("C")
    want.append("C " + enc([ref.squeeze_challenge()])[0].tobytes().hex())
    want.append("W " + ref.finalize().hex())
    assert py.finalize() == ref.finalize()
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    assert [l for l in out if l] == want


def _two_instances(k, seed):
    """a second satisfying witness for the same proving key: the fixture with another seed differs in the advice and
    instance columns and in rows of fixed column 5 beyond the usable range, which no constraint reads"""
    fx, oparams, opk, cs, eng, pk = both_sides(k, seed)
    fx2 = fxm.build(k=k, seed=seed + 100)
    usable = (1 << k) - cs.blinding_factors() - 1
    assert all(a[:usable] == b[:usable] for a, b in zip(fx["fixed"], fx2["fixed"])) and fx["mapping"] == fx2["mapping"]
    advs = [fx["advice"], fx2["advice"]]
    insts = [[fx["instance"][0][:4]], [fx2["instance"][0][:4]]]
    return oparams, opk, cs, eng, pk, advs, insts


@pytest.mark.parametrize("use_gwc", [True, False])
def test_multi_circuit_proof_bytes_match_oracle_and_verify(use_gwc):
    """create_proof_ext(circuits: &[C], instances: &[&[&[Fr]]]) (plonk/prover.rs:206-222): two instances of the
    circuit in one proof.  The oracle folds both through ONE evaluate_h accumulator as the reference does; the prover
    mirror evaluates each instance on its own and folds the h pieces with y^fold_steps -- the bytes must agree, the
    oracle verifier must accept them and must reject the instances in the other order."""
    k = 5
    oparams, opk, cs, eng, pk, advs, insts = _two_instances(k, 11)
    assert HP.fold_steps(cs) == PR.fold_steps(opk.vk.cs)
    want = PR.create_proof_multi(oparams, opk, advs, insts, HP.SeededRng(5), use_gwc=use_gwc)
    got = HP.create_proof_multi(HostParams(k), pk, [np.ascontiguousarray(np.stack([enc(c) for c in a])) for a in advs],
                                insts, HP.SeededRng(5), engine=eng, use_gwc=use_gwc)
    assert got == want
    assert PR.verify_proof_multi(oparams, opk.vk, insts, got, use_gwc=use_gwc)
    assert PR.verify_proof_multi(oparams, opk.vk, insts, got, use_gwc=use_gwc, pairing=True)
    assert not PR.verify_proof_multi(oparams, opk.vk, insts[::-1], got, use_gwc=use_gwc)
    single = PR.create_proof(oparams, opk, advs[0], insts[0], HP.SeededRng(5), use_gwc=use_gwc)
    assert len(got) > len(single)
    with pytest.raises(HP.B2Error):
        HP.create_proof_multi(HostParams(k), pk, [], [], HP.SeededRng(5), engine=eng)
    with pytest.raises(HP.B2Error):
        HP.create_proof_multi(HostParams(k), pk, [np.zeros((9, 32, 4), dtype=np.uint64)], insts, HP.SeededRng(5), engine=eng)
