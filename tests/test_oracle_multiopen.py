"""The reference's own multiopen tests (halo2_proofs/src/poly/multiopen.rs:209-411: test_roundtrip, test_multiopen),
restated for both arguments (GWC and SHPLONK): on the oracle (oracle/prover.py) and on the prover mirror's host logic
(halo2_gpu_specific_b200.plonk.gwc_create_proof / shplonk_create_proof over the oracle-backed engine), which must write
the oracle's bytes.  CPU only."""
import random

import pytest

from oracle import bn254 as o
from oracle import prover as PR
from oracle_engine import OracleEngine

from halo2_gpu_specific_b200 import plonk as HP
from halo2_gpu_specific_b200 import transcript as HT

R = o.R_MOD
enc = o.fr_encode
S_TOXIC = 0x2B200B200B200B200B200B200B200B2001


def _prove(params, queries, use_gwc):
    """oracle side: queries (rotation, point, poly as int list)"""
    tr = PR.Blake2bWrite()
    (PR.gwc_create_proof if use_gwc else PR.shplonk_create_proof)(params, tr, queries)
    return tr.finalize()


def _prove_mirror(params, k, queries, use_gwc):
    """prover mirror: the same queries with (n, 4) Montgomery arrays as polynomial handles"""
    eng = _engine(params, k)
    handles = {}
    qs = []
    for rot, point, poly in queries:
        h = handles.setdefault(id(poly), enc(poly))
        qs.append((rot, point, h))
    tr = HT.Blake2bWrite()
    (HP.gwc_create_proof if use_gwc else HP.shplonk_create_proof)(eng, tr, qs)
    return tr.finalize()


def _engine(params, k):
    eng = OracleEngine.__new__(OracleEngine)            # the multiopen provers need no constraint system
    eng.p, eng.d, eng.domain = params, o.EvaluationDomain(1, k), o.EvaluationDomain(1, k)
    return eng


def _verify(params, proof, vqueries, use_gwc, pairing=False):
    tr = PR.Blake2bRead(proof)
    left, right = (PR.gwc_verify_proof if use_gwc else PR.shplonk_verify_proof)(params, tr, vqueries)
    return PR.Decider.verify(params, left, right) if pairing else PR.Decider.verify_trapdoor(params, left, right)


@pytest.mark.parametrize("use_gwc", [True, False])
def test_roundtrip(use_gwc):
    """multiopen.rs:209-311"""
    k = 4
    n = 1 << k
    params = PR.Params(k, S_TOXIC)
    domain = o.EvaluationDomain(1, k)
    ax = [10 + i for i in range(n)]
    bx = [100 + i for i in range(n)]
    cx = [100 + i for i in range(n)]
    a, b, c = params.commit(ax), params.commit(bx), params.commit(cx)
    x = random.Random(1).randrange(R)
    y = domain.rotate_omega(x, 1)
    avx, bvx, cvy = o.eval_polynomial(ax, x), o.eval_polynomial(bx, x), o.eval_polynomial(cx, y)
    queries = [(0, x, ax), (0, x, bx), (1, y, cx)]
    proof = _prove(params, queries, use_gwc)
    assert _prove_mirror(params, k, queries, use_gwc) == proof
    one = lambda p: [(1, p)]                                                       # noqa: E731
    wrong = [(0, x, one(a), avx, "a"), (0, x, one(b), avx, "b"), (1, y, one(c), cvy, "c")]      # NB: wrong!
    assert not _verify(params, proof, wrong, use_gwc)
    right = [(0, x, one(a), avx, "a"), (0, x, one(b), bvx, "b"), (1, y, one(c), cvy, "c")]
    assert _verify(params, proof, right, use_gwc)
    assert _verify(params, proof, right, use_gwc, pairing=True)


@pytest.mark.parametrize("use_gwc", [True, False])
def test_multiopen(use_gwc):
    """multiopen.rs:313-411: nine rotation sets, set i opened by i polynomials"""
    k = 3
    n = 1 << k
    params = PR.Params(k, S_TOXIC)
    rng = random.Random(2)
    rotation_sets = [[1, 2, 3], [2, 3, 4], [2, 3, 4], [4, 5, 6, 7], [8], [9], [10, 11], [10, 11], [10, 11]]
    polys = [[[rng.randrange(R) for _ in range(n)] for _ in range(i)] for i in range(len(rotation_sets))]
    commitments = [[params.commit(p) for p in ps] for ps in polys]
    pq, vq = [], []
    for i, rots in enumerate(rotation_sets):
        for rot in rots:
            point = rot                                  # Fr::from(i as u64)
            for j in range(i):
                pq.append((rot, point, polys[i][j]))
                vq.append((rot, point, [(1, commitments[i][j])], o.eval_polynomial(polys[i][j], point), (i, j)))
    proof = _prove(params, pq, use_gwc)
    assert _prove_mirror(params, k, pq, use_gwc) == proof
    assert _verify(params, proof, vq, use_gwc)
    assert len(proof) == (32 * len({r for rs in rotation_sets[1:] for r in rs}) if use_gwc else 64)
    bad = list(vq)
    q = bad[7]
    bad[7] = (q[0], q[1], q[2], (q[3] + 1) % R, q[4])
    assert not _verify(params, proof, bad, use_gwc)
    swapped = bytearray(proof)
    swapped[5] ^= 2
    try:
        assert not _verify(params, bytes(swapped), vq, use_gwc)
    except (PR.TranscriptError, PR.VerifyError):
        pass


def test_intermediate_sets():
    """shplonk.rs:204-336 test_intermediate_sets: 100 random query lists; every evaluation comes back under its
    (commitment, point); the grouping does not depend on the point values; the sets cover exactly the queries;
    shuffling the queries changes neither property"""
    rng = random.Random(3)
    for _ in range(100):
        rotations = [(rng.randrange(R), rot) for rot in range(-4, 5)]            # (point, rotation)
        rotation_of = {p: r for p, r in rotations}
        evals = {(c, p): rng.randrange(R) for p, _ in rotations for c in range(8)}
        queries_0 = []
        for _q in range(16):
            c = rng.randrange(8)
            p, r = rng.choice(rotations)
            queries_0.append((r, p, c, evals[(c, p)]))

        def sets_of(queries):
            table = {(q[2], q[0]): q[3] for q in queries}
            return PR.shplonk_intermediate_sets(queries, lambda q: q[2], lambda k, r: table[(k, r)])[0]

        def check_evals(rotation_sets):
            for commitments, points in rotation_sets:
                for com, evs in commitments:
                    for e, p in zip(evs, points):
                        assert e == evals[(com, p)]

        def make_queries(rotation_sets):
            out = []
            for commitments, points in rotation_sets:
                for com, evs in commitments:
                    for p, e in zip(points, evs):
                        out.append((rotation_of[p], p, com, e))
            return out

        sets_0 = sets_of(queries_0)
        check_evals(sets_0)
        e = rng.randrange(R)
        moved = [(q[0], e * (q[0] % R) % R, q[2], q[3]) for q in queries_0]         # change points, keep rotations
        sets_1 = sets_of(moved)
        assert [[c for c in cs] for cs, _ in sets_0] == [[c for c in cs] for cs, _ in sets_1]
        rebuilt = make_queries(sets_0)
        assert set(rebuilt) == set(queries_0)
        shuffled = list(queries_0)
        rng.shuffle(shuffled)
        sets_2 = sets_of(shuffled)
        check_evals(sets_2)
        assert set(make_queries(sets_2)) == set(queries_0)
