"""Pins oracle/prover.py and oracle/pairing.py (CPU only) the way the reference pins its own prover: a proof
made by create_proof must be accepted by verify_proof, and must stop being accepted when the proof, the public
inputs or the witness change (examples/*.rs, plonk_api tests: create_proof -> verify_proof -> assert Ok / Err).
The verifier shares no prover code path beyond the transcript and the expression evaluator: it recomputes h(x)
from evaluations and checks openings against commitments, so prover and verifier pin each other."""
import hashlib

import pytest

from oracle import bn254 as o
from oracle import pairing as pr
from oracle import prover as PR

import plonk_fixture as fxm
from halo2_gpu_specific_b200.plonk import SeededRng

R = o.R_MOD
S_TOXIC = 0x2B200B200B200B200B200B200B200B2001


@pytest.fixture(scope="module")
def setup():
    k = 5
    fx = fxm.build(k=k, seed=11)
    params = PR.Params(k, S_TOXIC)
    pk = PR.keygen(params, fx["cs"], fx["fixed"], fx["mapping"])
    inst = [fx["instance"][0][:4]]
    proof = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1))
    return fx, params, pk, inst, proof


def test_proof_verifies(setup):
    fx, params, pk, inst, proof = setup
    cs = fx["cs"]
    q = pk.vk.queries
    # proof layout: commitments and evaluations in transcript order, 32 bytes each
    n_sets = 2
    points = cs.num_advice + len(cs.lookups) + n_sets + sum(len(l["input_expressions_sets"]) for l in cs.lookups) \
        + len(cs.shuffles) + 1 + pk.vk.domain.quotient_poly_degree
    evals = len(q["Instance"]) + len(q["Advice"]) + len(q["Fixed"]) + 1 + len(cs.permutation_columns) \
        + (3 * n_sets - 1) + sum(1 + 3 * len(l["input_expressions_sets"]) - 1 for l in cs.lookups) + 2 * len(cs.shuffles)
    rotations = {0, 1, -(cs.blinding_factors() + 1)} | {r for v in q.values() for _, r in v}
    assert len(proof) == 32 * (points + evals + len(rotations))
    assert PR.verify_proof(params, pk.vk, inst, proof)


def test_proof_verifies_with_pairing(setup):
    """Decider::verify as the reference runs it: e(left, [s]G2) * e(right, -G2) == 1, no toxic waste in G1"""
    fx, params, pk, inst, proof = setup
    assert PR.verify_proof(params, pk.vk, inst, proof, pairing=True)
    bad = bytearray(proof)
    bad[40] ^= 1
    try:
        assert not PR.verify_proof(params, pk.vk, inst, bytes(bad), pairing=True)
    except (PR.TranscriptError, PR.VerifyError):
        pass


def test_every_proof_element_is_bound(setup):
    """flipping one bit in any 32-byte element (commitment, evaluation or opening witness) must be caught"""
    fx, params, pk, inst, proof = setup
    for el in range(len(proof) // 32):
        bad = bytearray(proof)
        bad[32 * el + 3] ^= 0x10
        try:
            ok = PR.verify_proof(params, pk.vk, inst, bytes(bad))
        except (PR.TranscriptError, PR.VerifyError):
            continue                      # not a curve point / not canonical: rejected while reading
        assert not ok, f"element {el} is not bound by the verifier"


def test_wrong_public_input_is_rejected(setup):
    fx, params, pk, inst, proof = setup
    bad = [[(inst[0][0] + 1) % R] + list(inst[0][1:])]
    assert not PR.verify_proof(params, pk.vk, bad, proof)
    assert not PR.verify_proof(params, pk.vk, [inst[0][:3]], proof)


@pytest.mark.parametrize("what", ["witness", "copy", "shuffle"])
def test_unsatisfied_circuit_is_rejected(what):
    k = 5
    fx = fxm.build(k=k, seed=11, break_what=what)
    params = PR.Params(k, S_TOXIC)
    pk = PR.keygen(params, fx["cs"], fx["fixed"], fx["mapping"])
    inst = [fx["instance"][0][:4]]
    proof = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1))
    assert not PR.verify_proof(params, pk.vk, inst, proof)


def test_lookup_outside_table_cannot_be_proven():
    fx = fxm.build(k=5, seed=11)
    fx["advice"][7][3] = 1                # not a table value (table = i^2 + 3)
    params = PR.Params(5, S_TOXIC)
    pk = PR.keygen(params, fx["cs"], fx["fixed"], fx["mapping"])
    with pytest.raises(KeyError):         # .expect("logup binary_search_by_key should hit"), logup/prover.rs:148
        PR.create_proof(params, pk, fx["advice"], [fx["instance"][0][:4]], SeededRng(1))


def test_proof_is_deterministic_under_a_fixed_rng_and_frozen(setup):
    """north star: proof transcript bytes under a fixed RNG.  The digest is frozen so that any drift in the
    oracle (transcript order, RNG order, encoding) is visible."""
    fx, params, pk, inst, proof = setup
    again = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1))
    assert again == proof
    other = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(2))
    assert other != proof and PR.verify_proof(params, pk.vk, inst, other)
    import json
    import os
    golden = os.path.join(os.path.dirname(__file__), "golden", "proof_digest.json")
    want = json.load(open(golden))
    assert hashlib.sha256(proof).hexdigest() == want["k5_seed11_rng1_sha256"]
    shplonk = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1), use_gwc=False)
    assert hashlib.sha256(shplonk).hexdigest() == want["k5_seed11_rng1_shplonk_sha256"]


def test_multiplicity_tie_rule():
    """table with repeated values: the row binary_search_by_key lands on gets the whole count"""
    table = [5, 7, 7, 7, 9, 9, 11, 3]
    usable = 8
    m = PR.logup_multiplicity([[[7, 9, 7, 3, 11, 5, 9, 9]]], table, usable, 8)
    assert sum(m) == 8
    pairs = sorted(((table[i], i) for i in range(usable)), key=lambda p: p[0])
    for v in set(table):
        row = PR.binary_search_index(pairs, v)
        assert table[row] == v
        assert m[row] == [7, 9, 7, 3, 11, 5, 9, 9].count(v)


def test_pairing_is_bilinear_and_non_degenerate():
    a = 0x1234567890ABCDEF1234567
    assert pr.g2_is_on_curve(pr.G2_GEN)
    assert pr.g2_add(pr.g2_mul(pr.G2_GEN, R - 1), pr.G2_GEN) is None
    e0 = pr.pairing(o.G1_GEN, pr.G2_GEN)
    e1 = pr.pairing(o.g1_mul(o.G1_GEN, a), pr.G2_GEN)
    e2 = pr.pairing(o.G1_GEN, pr.g2_mul(pr.G2_GEN, a))
    assert e0 != pr.F12_ONE and e1 == e2 == pr.f12_pow(e0, a)
    assert pr.f12_pow(e0, R) == pr.F12_ONE
    assert pr.pairing_check([(o.g1_mul(o.G1_GEN, a), pr.G2_GEN), (o.g1_neg(o.G1_GEN), pr.g2_mul(pr.G2_GEN, a))])
    assert not pr.pairing_check([(o.g1_mul(o.G1_GEN, a + 1), pr.G2_GEN), (o.g1_neg(o.G1_GEN), pr.g2_mul(pr.G2_GEN, a))])


def test_transcript_known_answers():
    """Blake2b, personal "Halo2-Transcript", 64-byte digest, prefix bytes 0 / 1 / 2, challenge = digest mod r
    (transcript.rs:14-20, 121-147, 266-276) against an independent hashlib computation"""
    tr = PR.Blake2bWrite()
    tr.common_scalar(5)
    tr.write_point((1, 2))
    c = tr.squeeze_challenge()
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
    h.update(b"\x02" + (5).to_bytes(32, "little"))
    h.update(b"\x01" + (1).to_bytes(32, "little") + (2).to_bytes(32, "little"))
    h.update(b"\x00")
    assert c == int.from_bytes(h.digest(), "little") % R
    assert tr.finalize() == o.g1_to_bytes((1, 2))
    c2 = tr.squeeze_challenge()          # the state keeps absorbing: a second squeeze differs
    assert c2 != c
    rd = PR.Blake2bRead(tr.finalize())
    rd.common_scalar(5)
    assert rd.read_point() == (1, 2)
    assert rd.squeeze_challenge() == c


def test_shplonk_proof_verifies_and_binds_every_element(setup):
    """create_proof_with_shplonk + the SHPLONK verifier (poly/multiopen/shplonk/{prover,verifier}.rs)"""
    fx, params, pk, inst, _ = setup
    proof = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1), use_gwc=False)
    assert PR.verify_proof(params, pk.vk, inst, proof, use_gwc=False)
    assert PR.verify_proof(params, pk.vk, inst, proof, use_gwc=False, pairing=True)
    for el in range(len(proof) // 32):
        bad = bytearray(proof)
        bad[32 * el + 3] ^= 0x10
        try:
            ok = PR.verify_proof(params, pk.vk, inst, bytes(bad), use_gwc=False)
        except (PR.TranscriptError, PR.VerifyError):
            continue
        assert not ok, f"element {el} is not bound by the SHPLONK verifier"
    assert not PR.verify_proof(params, pk.vk, [[(inst[0][0] + 1) % R] + list(inst[0][1:])], proof, use_gwc=False)
    # a GWC proof is not a SHPLONK proof
    gwc = PR.create_proof(params, pk, fx["advice"], inst, SeededRng(1))
    try:
        assert not PR.verify_proof(params, pk.vk, inst, gwc, use_gwc=False)
    except (PR.TranscriptError, PR.VerifyError):
        pass


def test_shplonk_intermediate_sets_grouping():
    """shplonk.rs:57-150: commitments grouped by their SET of rotations, sets in BTreeSet order, first-appearance
    order inside a set, super point set ordered by rotation"""
    pts = {-1: 11, 0: 22, 1: 33}
    q = lambda r, c: (r, pts[r], c, 100 * r + ord(c))                                     # noqa: E731
    queries = [q(0, "a"), q(1, "b"), q(0, "b"), q(-1, "c"), q(0, "c"), q(0, "d"), q(1, "e"), q(0, "e")]
    sets, super_points = PR.shplonk_intermediate_sets(queries, lambda x: x[2], lambda k, r: 100 * r + ord(k))
    assert super_points == [11, 22, 33]
    assert [([c for c, _ in cs], p) for cs, p in sets] == [(["c"], [11, 22]), (["a", "d"], [22]), (["b", "e"], [22, 33])]
    assert sets[2][0][0][1] == [ord("b"), 100 + ord("b")]
