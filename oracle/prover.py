"""CPU oracle for the whole proof: keygen, create_proof and verify_proof (GWC multiopen, the `create_proof` default).

TEST INFRASTRUCTURE ONLY (same rules as oracle/bn254.py and oracle/plonk.py): only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Restates, with Python big ints (MSMs go through the C restatement, oracle/cpu_ref.c):
  * transcript::{Blake2bWrite, Blake2bRead, Challenge255}        transcript.rs:14-293
  * VerifyingKey::hash_into                                     plonk.rs:91-109   (see `vk_transcript_repr`)
  * keygen_vk / keygen_pk (fixed + permutation parts)           plonk/keygen.rs:213-431, plonk/permutation/keygen.rs:196-262
  * create_single_instances / create_proof_from_witness         plonk/prover.rs:85-173, 916-1500 (same body as :206-850)
  * vanishing::Argument::{commit, construct, evaluate, open}    plonk/vanishing/prover.rs:41-153
  * permutation / logup / shuffle evaluate + open               plonk/permutation/prover.rs:181-304, plonk/logup/prover.rs:420-491,
                                                                plonk/shuffle/prover.rs:200-240
  * multiopen::gwc::{create_proof, verify_proof}                poly/multiopen/gwc.rs:38-62, gwc/prover.rs:19-173, gwc/verifier.rs:16-91
  * multiopen::shplonk::{create_proof, verify_proof}            poly/multiopen/shplonk.rs:57-150, shplonk/prover.rs:78-234,
                                                                shplonk/verifier.rs:23-104 (PreMSM / combine_with_base: poly/msm.rs:136-204)
  * verify_proof                                                plonk/verifier.rs:127-507 and the argument verifiers
                                                                (vanishing/verifier.rs, permutation/verifier.rs,
                                                                logup/verifier.rs, shuffle/verifier.rs)
  * Decider::verify                                             poly/multiopen.rs:31-57
All paths relative to /root/reference/halo2_proofs/src.

PARITY STATUS: byte-level parity with the Rust binary is unpinned (the reference cannot be built here and holds
no golden proofs).  What pins this file is the reference's own acceptance test strategy (examples / tests:
create_proof -> verify_proof must accept, a tampered proof / instance must not): tests/test_oracle_prover.py, and the
reference's multiopen unit tests restated in tests/test_oracle_multiopen.py.
Three inputs are parameters because they cannot be read from the reference tree:
  * the verifying key's transcript scalar (Rust `{:?}` of PinnedVerificationKey, plonk.rs:100) -- `vk_transcript_repr`
    hashes a description of its own under the same personalisation and feeds it through the same common_scalar call;
  * the point compression rule ([EXT], oracle/bn254.py g1_to_bytes) -- `sign_bit`;
  * randomness: the reference draws from OsRng / thread_rng at four sites (plonk/prover.rs:283, 352, 427;
    vanishing/prover.rs:57); here every draw comes from the caller's `rng` in the order `create_proof` documents
    ("thread the caller's rng through").
The pairing check e(left, [s]G2) * e(right, -G2) == 1 is restated in two equivalent forms: `Decider.verify_trapdoor`
checks [s]*left == right in G1 with the toxic waste of the synthetic SRS (Params::unsafe_setup keeps no s, the
test SRS does), and `Decider.verify` runs the optimal-ate pairing (oracle/pairing.py) on [s]G2 alone.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import bn254 as o
from . import cref
from . import plonk as P

R = o.R_MOD
Q = o.Q_MOD
Point = Optional[Tuple[int, int]]


# --------------------------------------------------------------------------
# transcript.rs
# --------------------------------------------------------------------------
PREFIX_CHALLENGE, PREFIX_POINT, PREFIX_SCALAR = b"\x00", b"\x01", b"\x02"     # transcript.rs:14-20


class TranscriptError(Exception):
    pass


class _Blake2b:
    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")     # :83-86, :161-164

    def squeeze_challenge(self) -> int:
        """:121-126, Challenge255::new :266-276 (from_bytes_wide of the 64-byte digest)"""
        self.state.update(PREFIX_CHALLENGE)
        return int.from_bytes(self.state.copy().digest(), "little") % R

    def common_point(self, p: Point) -> None:
        """:128-140"""
        if p is None:
            raise TranscriptError("cannot write points at infinity to the transcript")
        self.state.update(PREFIX_POINT)
        self.state.update(p[0].to_bytes(32, "little"))
        self.state.update(p[1].to_bytes(32, "little"))

    def common_scalar(self, s: int) -> None:
        """:142-147"""
        self.state.update(PREFIX_SCALAR)
        self.state.update((s % R).to_bytes(32, "little"))


class Blake2bWrite(_Blake2b):
    def __init__(self, sign_bit: int = 7):
        super().__init__()
        self.sign_bit = sign_bit
        self.proof = bytearray()

    def write_point(self, p: Point) -> None:
        """:180-184"""
        self.common_point(p)
        self.proof += o.g1_to_bytes(p, self.sign_bit)

    def write_scalar(self, s: int) -> None:
        """:185-189"""
        self.common_scalar(s)
        self.proof += (s % R).to_bytes(32, "little")

    def finalize(self) -> bytes:
        return bytes(self.proof)


class Blake2bRead(_Blake2b):
    def __init__(self, proof: bytes, sign_bit: int = 7):
        super().__init__()
        self.sign_bit = sign_bit
        self.data = bytes(proof)
        self.pos = 0

    def _take(self) -> bytes:
        if self.pos + 32 > len(self.data):
            raise TranscriptError("proof too short")
        b = self.data[self.pos:self.pos + 32]
        self.pos += 32
        return b

    def read_point(self) -> Point:
        """:93-102"""
        try:
            p = o.g1_from_bytes(self._take(), self.sign_bit)
        except ValueError as e:
            raise TranscriptError(f"invalid point encoding in proof: {e}")
        self.common_point(p)
        return p

    def read_scalar(self) -> int:
        """:104-116"""
        s = int.from_bytes(self._take(), "little")
        if s >= R:
            raise TranscriptError("invalid field element encoding in proof")
        self.common_scalar(s)
        return s


# --------------------------------------------------------------------------
# ConstraintSystem queries (plonk/circuit.rs query_*_index / get_any_query_index)
# --------------------------------------------------------------------------
def walk_expression(e, visit) -> None:
    t = e[0]
    if t in ("Fixed", "Advice", "Instance"):
        visit(t, e[1], e[2])
    elif t in ("Negated", "Scaled"):
        walk_expression(e[1], visit)
    elif t in ("Sum", "Product"):
        walk_expression(e[1], visit)
        walk_expression(e[2], visit)


def collect_queries(cs) -> Dict[str, List[Tuple[int, int]]]:
    """The (column, rotation) lists ConstraintSystem keeps per column kind.  In the reference their order is the
    order of the circuit's meta.query_* calls (a front-end matter); the rule here: permutation columns at
    Rotation::cur first (enable_equality, circuit.rs query_any_index), then gates, lookups (input sets, table),
    shuffles, each in order of first appearance.  A cs that carries explicit *_queries lists keeps them."""
    if getattr(cs, "advice_queries", None) is not None:
        return {"Advice": cs.advice_queries, "Fixed": cs.fixed_queries, "Instance": cs.instance_queries}
    q: Dict[str, List[Tuple[int, int]]] = {"Advice": [], "Fixed": [], "Instance": []}

    def visit(kind, col, rot):
        if (col, rot) not in q[kind]:
            q[kind].append((col, rot))

    for kind, col in cs.permutation_columns:
        visit(kind, col, 0)
    for gate in cs.gates:
        for poly in gate:
            walk_expression(poly, visit)
    for lk in cs.lookups:
        for s in lk["input_expressions_sets"]:
            for inp in s:
                for e in inp:
                    walk_expression(e, visit)
        for e in lk["table_expressions"]:
            walk_expression(e, visit)
    for group in cs.shuffles:
        for a in group:
            for e in a["input_expressions"] + a["shuffle_expressions"]:
                walk_expression(e, visit)
    cs.advice_queries, cs.fixed_queries, cs.instance_queries = q["Advice"], q["Fixed"], q["Instance"]
    return q


def eval_expression_at_queries(e, queries, fixed_evals, advice_evals, instance_evals) -> int:
    """Expression::evaluate with the verifier's closures (plonk/verifier.rs:318-331): a column query reads the
    evaluation at its query index."""
    t = e[0]
    if t == "Constant":
        return e[1]
    if t in ("Fixed", "Advice", "Instance"):
        evals = {"Fixed": fixed_evals, "Advice": advice_evals, "Instance": instance_evals}[t]
        return evals[queries[t].index((e[1], e[2]))]
    rec = lambda x: eval_expression_at_queries(x, queries, fixed_evals, advice_evals, instance_evals)  # noqa: E731
    if t == "Negated":
        return (-rec(e[1])) % R
    if t == "Sum":
        return (rec(e[1]) + rec(e[2])) % R
    if t == "Product":
        return rec(e[1]) * rec(e[2]) % R
    if t == "Scaled":
        return rec(e[1]) * e[2] % R
    raise ValueError(t)


# --------------------------------------------------------------------------
# Params (synthetic SRS through the C restatement) and commitments
# --------------------------------------------------------------------------
def _canonical_limbs(vals: Sequence[int]) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = o._to_limbs(v % R)
    return out


def msm(scalars: Sequence[int], bases: np.ndarray) -> Point:
    """best_multiexp (arithmetic.rs:465-492, C restatement) -> affine point"""
    if len(scalars) == 0:
        return None
    j = cref.best_multiexp(o.fr_encode(scalars), bases[:len(scalars)])
    return o.g1_affine_decode(cref.jac_to_affine(j))[0]


class Params:
    """poly/commitment.rs:56-124 (unsafe_setup) with a caller-chosen s; bases as (n, 8) Montgomery affine arrays
    (the layout the engine registers), commit / commit_lagrange :129-142."""

    def __init__(self, k: int, s: int):
        self.k, self.n, self.s = k, 1 << k, s % R
        n = self.n
        pw, cur = [], 1
        for _ in range(n):
            pw.append(cur)
            cur = cur * s % R
        self.g = cref.g1_mul_gen(_canonical_limbs(pw))                                # :63-83
        root = o.FR_ROOT_OF_UNITY
        for _ in range(k, o.FR_S):
            root = root * root % R
        mult = (pow(s, n, R) - 1) * o.fr_inv(n % R) % R
        lag, rp = [], 1
        for _ in range(n):
            lag.append(mult * rp % R * o.fr_inv((s - rp) % R) % R)                    # :85-112
            rp = rp * root % R
        self.g_lagrange = cref.g1_mul_gen(_canonical_limbs(lag))
        self.g1 = o.G1_GEN                                                            # ParamsVerifier.g1

    def commit(self, poly: Sequence[int]) -> Point:
        assert len(poly) <= self.n
        return msm(poly, self.g)

    def commit_lagrange(self, poly: Sequence[int]) -> Point:
        assert len(poly) <= self.n
        return msm(poly, self.g_lagrange)


class ParamsVerifier:
    """What verify_proof reads from the parameters (poly/commitment.rs:296-389 ParamsVerifier: g1, s_g2, g_lagrange
    for the public inputs) without materialising 2 * 2^k points: g_lagrange[i] is computed on demand, so proofs
    over large domains can be checked when the SRS itself was built on the device (Params.unsafe_setup)."""

    def __init__(self, k: int, s: int):
        self.k, self.n, self.s = k, 1 << k, s % R
        self.g1 = o.G1_GEN
        root = o.FR_ROOT_OF_UNITY
        for _ in range(k, o.FR_S):
            root = root * root % R
        self._root = root
        self._mult = (pow(s, self.n, R) - 1) * o.fr_inv(self.n % R) % R

    def commit_lagrange(self, poly: Sequence[int]) -> Point:
        acc: Point = None
        for i, v in enumerate(poly):
            if v % R:
                rp = pow(self._root, i, R)
                li = self._mult * rp % R * o.fr_inv((self.s - rp) % R) % R          # commitment.rs:85-112
                acc = o.g1_add(acc, o.g1_mul(o.G1_GEN, li * v % R))
        return acc


# --------------------------------------------------------------------------
# keygen
# --------------------------------------------------------------------------
def vk_transcript_repr(cs, domain, fixed_commitments, permutation_commitments) -> int:
    """plonk.rs:91-109 hashes Rust's `{:?}` of the pinned verifying key (Blake2b, personal "Halo2-Verify-Key",
    u64 length prefix) and absorbs from_bytes_wide(digest) with common_scalar.  The Debug text is not
    reproducible without the crate, so the string hashed here is this file's own description of the same
    contents; callers that hold the reference's scalar pass it to create_proof / verify_proof instead."""
    q = collect_queries(cs)
    s = repr({
        "base_modulus": hex(Q), "scalar_modulus": hex(R),
        "domain": {"k": domain.k, "extended_k": domain.extended_k, "omega": hex(domain.omega)},
        "fixed_commitments": fixed_commitments, "permutation": permutation_commitments,
        "cs": {"num_fixed": cs.num_fixed, "num_advice": cs.num_advice, "num_instance": cs.num_instance,
               "gates": cs.gates, "lookups": cs.lookups, "shuffles": cs.shuffles,
               "permutation_columns": cs.permutation_columns, "queries": q,
               "degree": cs.degree(), "blinding_factors": cs.blinding_factors()},
    }).encode()
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(len(s).to_bytes(8, "little"))
    h.update(s)
    return int.from_bytes(h.digest(), "little") % R


class VerifyingKey:
    def __init__(self, cs, domain, fixed_commitments, permutation_commitments, transcript_repr=None):
        self.cs, self.domain = cs, domain
        self.fixed_commitments = fixed_commitments
        self.permutation_commitments = permutation_commitments
        self.queries = collect_queries(cs)
        self.transcript_repr = (vk_transcript_repr(cs, domain, fixed_commitments, permutation_commitments)
                                if transcript_repr is None else transcript_repr % R)


class ProvingKey:
    pass


def keygen(params: Params, cs, fixed: Sequence[Sequence[int]], mapping, zeta: Optional[int] = None,
           transcript_repr: Optional[int] = None) -> ProvingKey:
    """keygen_vk + keygen_pk (plonk/keygen.rs:213-431) for an already laid out circuit: `fixed` are the fixed
    columns (selectors included, already compressed or not -- a front-end matter), `mapping` the permutation
    (oracle/plonk.py identity_mapping / mapping_copy)."""
    domain = o.EvaluationDomain(cs.degree(), params.k) if zeta is None else o.EvaluationDomain(cs.degree(), params.k, zeta)
    assert len(fixed) == cs.num_fixed and all(len(c) == params.n for c in fixed)
    sigmas = P.permutation_sigmas(cs, domain, mapping)
    fixed_commitments = [params.commit_lagrange(c) for c in fixed]               # keygen.rs:288-291
    permutation_commitments = [params.commit_lagrange(s) for s in sigmas]        # permutation/keygen.rs:244-251
    pk = ProvingKey()
    pk.vk = VerifyingKey(cs, domain, fixed_commitments, permutation_commitments, transcript_repr)
    pk.fixed_values = [list(c) for c in fixed]
    pk.fixed_polys = [domain.lagrange_to_coeff(c) for c in fixed]               # keygen.rs:381-385
    pk.sigmas = sigmas
    pk.sigma_polys = [domain.lagrange_to_coeff(s) for s in sigmas]              # permutation/keygen.rs:254-262
    pk.l0, pk.l_last, pk.l_active_row = P.lagrange_basis_cosets(cs, domain)     # keygen.rs:398-431
    pk.ev = P.Evaluator.new(cs)                                                  # keygen.rs:433
    return pk


# --------------------------------------------------------------------------
# randomness: adapter from the caller's vector RNG to the scalar draws of oracle/plonk.py
# --------------------------------------------------------------------------
class _RngAdapter:
    """oracle/plonk.py draws blinding values one at a time with rng.randrange(modulus); the proof-level RNG
    contract is in blocks (create_proof docstring).  A block is fetched when the first value of it is asked for."""

    def __init__(self, rng, fr_block: int, u16_block: int):
        self.rng, self.fr_block, self.u16_block = rng, fr_block, u16_block
        self.fr_q: List[int] = []
        self.u16_q: List[int] = []

    def randrange(self, m: int) -> int:
        if m == R:
            if not self.fr_q:
                self.fr_q = o.fr_decode(self.rng.fr_vec(self.fr_block))
            return self.fr_q.pop(0)
        if m == 1 << 16:
            if not self.u16_q:
                self.u16_q = [int(v) for v in self.rng.u16_vec(self.u16_block)]
            return self.u16_q.pop(0)
        raise ValueError(m)


# --------------------------------------------------------------------------
# create_proof
# --------------------------------------------------------------------------


def chacha20_block(key: bytes, counter: int, nonce=(0, 0, 0)) -> bytes:
    """RFC 8439 section 2.3: one 64-byte key stream block (scalar restatement; the product's vectorised one is
    halo2_gpu_specific_b200/plonk.py chacha20_blocks, the device's csrc/chacha.cuh)"""
    M = 0xFFFFFFFF
    rotl = lambda v, c: ((v << c) & M) | (v >> (32 - c))                       # noqa: E731
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + \
        [int.from_bytes(key[4 * i:4 * i + 4], "little") for i in range(8)] + [counter & M] + [w & M for w in nonce]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 16)                 # noqa: E702
        x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 12)                 # noqa: E702
        x[a] = (x[a] + x[b]) & M; x[d] = rotl(x[d] ^ x[a], 8)                  # noqa: E702
        x[c] = (x[c] + x[d]) & M; x[b] = rotl(x[b] ^ x[c], 7)                  # noqa: E702

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)   # noqa: E702
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)   # noqa: E702
    return b"".join(((a + b) & M).to_bytes(4, "little") for a, b in zip(x, s))


def vanishing_random_poly(domain, rng) -> List[int]:
    """vanishing/prover.rs:48-63 with the caller's rng: random = k field elements (fr_vec(k));
    coeff[i] = (a_i + random[u_i % k]) * (b_i + random[v_i % k]).  The reference draws a_i, u_i, b_i, v_i from
    thread_rng (ChaCha under a 256-bit key) per coefficient; here they come from the ChaCha20 key stream (RFC 8439,
    nonce 0) under the 256-bit key u64_vec(4) of the rng: a_i = block 3i as a 512-bit little-endian integer mod r,
    b_i = block 3i + 1 likewise, u_i and v_i the first two little-endian 64-bit words of block 3i + 2."""
    k, n = domain.k, domain.n
    random = o.fr_decode(rng.fr_vec(k))
    key = b"".join(int(w).to_bytes(8, "little") for w in rng.u64_vec(4))
    out = []
    for i in range(n):
        a = int.from_bytes(chacha20_block(key, 3 * i), "little") % R
        b = int.from_bytes(chacha20_block(key, 3 * i + 1), "little") % R
        third = chacha20_block(key, 3 * i + 2)
        u, v = int.from_bytes(third[:8], "little"), int.from_bytes(third[8:16], "little")
        out.append((a + random[u % k]) * (b + random[v % k]) % R)
    return out


def scalar_bits(v: int) -> int:
    return v.bit_length()


def binary_search_index(sorted_pairs, value: int) -> int:
    """<[T]>::binary_search_by_key as implemented by the pinned toolchain (nightly-2023-06-01, core::slice):
    probes mid = left + size / 2 and returns the first probe that compares Equal -- which decides the row that
    gets the multiplicity when the table repeats a value (logup/prover.rs:115-117, 146-150)."""
    size = len(sorted_pairs)
    left, right = 0, size
    while left < right:
        mid = left + size // 2
        t = sorted_pairs[mid][0]
        if t < value:
            left = mid + 1
        elif t > value:
            right = mid
        else:
            return sorted_pairs[mid][1]
        size = right - left
    raise KeyError("logup binary_search_by_key should hit")


def logup_multiplicity(input_sets, table, usable: int, n: int) -> List[int]:
    """logup/prover.rs:115-184: stable sort of (value, row) by value (canonical integer order, Ord for Fr is
    [EXT]), binary search per input value, counts credited to the row the search returns."""
    pairs = sorted(((table[i], i) for i in range(usable)), key=lambda p: p[0])
    m = [0] * n
    cache: Dict[int, int] = {}
    for s in input_sets:
        for inp in s:
            for v in inp[:usable]:
                if v not in cache:
                    cache[v] = binary_search_index(pairs, v)
                m[cache[v]] += 1
    return m


def create_proof(params: Params, pk: ProvingKey, advice: Sequence[Sequence[int]], instances: Sequence[Sequence[int]],
                 rng, sign_bit: int = 7, use_gwc: bool = True) -> bytes:
    """one circuit instance per proof: create_proof_multi with lists of one"""
    return create_proof_multi(params, pk, [advice], [instances], rng, sign_bit, use_gwc)


def fold_steps(cs) -> int:
    """how many times evaluate_h multiplies its accumulator by y for ONE circuit instance (evaluation.rs:839-1220):
    every gate polynomial, the permutation terms, the lookup terms, the shuffle terms"""
    t = sum(len(g) for g in cs.gates)
    if cs.permutation_columns:
        chunk_len = cs.degree() - 2
        sets = (len(cs.permutation_columns) + chunk_len - 1) // chunk_len
        t += 2 + (sets - 1) + sets
    for lk in cs.lookups:
        t += 3 + 2 * (len(lk["input_expressions_sets"]) - 1)
    return t + 3 * len(cs.shuffles)


def create_proof_multi(params: Params, pk: ProvingKey, advices: Sequence[Sequence[Sequence[int]]],
                       instances: Sequence[Sequence[Sequence[int]]], rng, sign_bit: int = 7,
                       use_gwc: bool = True) -> bytes:
    """plonk/prover.rs:916-1500 (create_proof_from_witness: the advice columns are given, as read by fetch_witness)
    with the GWC multiopen (`create_proof`, :1759-1781: use_gwc = true) or SHPLONK (`create_proof_with_shplonk`,
    :1737-1757: use_gwc = false), for `len(advices)` instances of the same circuit in ONE proof
    (`circuits: &[ConcreteCircuit], instances: &[&[&[C::Scalar]]]`, plonk/prover.rs:206-222): every phase loops over
    the circuit instances before the next challenge is drawn, and evaluate_h folds them into one h(X)
    (evaluation.rs:839-845).

    `rng` supplies every random value, in this order (vector draws; fr_vec returns Montgomery limbs):
      1. per circuit: u16_vec(num_advice * (bf + 1)): blinding rows of advice column i are [i*(bf+1), (i+1)*(bf+1))   (:973-977)
      2. per circuit, per lookup: u16_vec(bf + 1) for the blinding rows of m                            (logup/prover.rs:232-236)
      3. per circuit, per permutation set: fr_vec(bf)                                                   (permutation/prover.rs:156-158)
      4. per circuit, per lookup, per z: fr_vec(bf)                                                     (plonk/prover.rs:445-449)
      5. per circuit, per shuffle group: fr_vec(bf)                                                     (plonk/prover.rs:518-521)
      6. vanishing_random_poly: fr_vec(k), u64_vec(4)                                                   (vanishing/prover.rs:48-63)
    """
    vk = pk.vk
    cs, domain = vk.cs, vk.domain
    n, k = params.n, params.k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    queries = vk.queries
    tr = Blake2bWrite(sign_bit)
    lag2coeff = domain.lagrange_to_coeff
    C = len(advices)
    if len(instances) != C:
        raise ValueError("InvalidInstances")

    # ---- create_single_instances :85-173
    tr.common_scalar(vk.transcript_repr)                                       # vk.hash_into
    inst_values, inst_polys = [], []
    for circuit_instances in instances:
        if len(circuit_instances) != cs.num_instance:
            raise ValueError("InvalidInstances")
        vals = []
        for values in circuit_instances:
            if len(values) > usable:
                raise ValueError("InstanceTooLarge")
            vals.append([v % R for v in values] + [0] * (n - len(values)))
        for poly in vals:
            tr.common_point(params.commit_lagrange(poly))                      # :124-137
        inst_values.append(vals)
        inst_polys.append([lag2coeff(p) for p in vals])

    # ---- advice :964-1010
    adv_values = []
    for advice in advices:
        assert len(advice) == cs.num_advice
        blind = [int(v) for v in rng.u16_vec(cs.num_advice * (bf + 1))]
        cols = []
        for i, col in enumerate(advice):
            col = [v % R for v in col[:usable]] + blind[i * (bf + 1):(i + 1) * (bf + 1)]
            assert len(col) == n
            cols.append(col)
        for col in cols:
            tr.write_point(params.commit_lagrange(col))                        # commit_lagrange_with_bound: same point
        adv_values.append(cols)
    theta = tr.squeeze_challenge()

    # ---- lookups: compress, m commitments (:334-366)
    adapter = _RngAdapter(rng, bf, bf + 1)
    all_lookups = []
    for advice_values, instance_values in zip(adv_values, inst_values):
        lookups = []
        for lk in cs.lookups:
            comp = lambda ex: P.evaluate_with_theta(ex, n, 1, pk.fixed_values, advice_values, instance_values, theta)  # noqa: E731
            input_sets = [[comp(inp) for inp in s] for s in lk["input_expressions_sets"]]
            table = comp(lk["table_expressions"])
            m = logup_multiplicity(input_sets, table, usable, n)
            for i in range(usable, n):
                m[i] = adapter.randrange(1 << 16)
            lookups.append({"input_sets": input_sets, "table": table, "m": m})
        all_lookups.append(lookups)
    for lookups in all_lookups:
        for lk in lookups:
            tr.write_point(params.commit_lagrange(lk["m"]))
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()

    # ---- z columns (:411-633); transcript order: permutations of every circuit, lookups of every circuit, shuffles
    all_perm_z = [P.permutation_commit(cs, domain, pk.sigmas, av, pk.fixed_values, iv, beta, gamma, adapter)
                  if cs.permutation_columns else [] for av, iv in zip(adv_values, inst_values)]
    for lookups in all_lookups:
        for lk in lookups:
            zs = P.logup_commit_z(cs, domain, lk["input_sets"], lk["table"], lk["m"], beta)
            lk["z"] = [P.blind_to_n(z, n, adapter) for z in zs]
    all_shuffle_z = [[P.blind_to_n(P.shuffle_commit_product(cs, domain, g, theta, beta, av, pk.fixed_values, iv), n, adapter)
                      for g in cs.shuffles] for av, iv in zip(adv_values, inst_values)]
    for perm_z in all_perm_z:
        for z in perm_z:
            tr.write_point(params.commit_lagrange(z))
    for lookups in all_lookups:
        for lk in lookups:
            for z in lk["z"]:
                tr.write_point(params.commit_lagrange(z))
    for shuffle_z in all_shuffle_z:
        for z in shuffle_z:
            tr.write_point(params.commit_lagrange(z))
    all_perm_polys = [[lag2coeff(z) for z in perm_z] for perm_z in all_perm_z]
    for lookups in all_lookups:
        for lk in lookups:
            lk["z_polys"] = [lag2coeff(z) for z in lk["z"]]
            lk["m_poly"] = lag2coeff(lk["m"])
    all_shuffle_polys = [[lag2coeff(z) for z in shuffle_z] for shuffle_z in all_shuffle_z]

    # ---- vanishing commit, y (:635-639)
    random_poly = vanishing_random_poly(domain, rng)
    tr.write_point(params.commit(random_poly))
    y = tr.squeeze_challenge()

    # ---- h(X) (:640-690, vanishing/prover.rs:64-110): one accumulator folded through every circuit instance
    adv_polys = [[lag2coeff(c) for c in cols] for cols in adv_values]
    ext = domain.coeff_to_extended
    fixed_cosets, sigma_cosets = [ext(p) for p in pk.fixed_polys], [ext(p) for p in pk.sigma_polys]
    h = None
    for ci in range(C):
        lookups = all_lookups[ci]
        h = P.evaluate_h(pk.ev, cs, domain, fixed_cosets, [ext(p) for p in adv_polys[ci]],
                         [ext(p) for p in inst_polys[ci]], pk.l0, pk.l_last, pk.l_active_row,
                         sigma_cosets, y, beta, gamma, theta,
                         [{"z_cosets": [ext(z) for z in lk["z_polys"]], "m_coset": ext(lk["m_poly"])} for lk in lookups],
                         [ext(p) for p in all_shuffle_polys[ci]], [ext(p) for p in all_perm_polys[ci]], values=h)
    h_coeffs = domain.extended_to_coeff(domain.divide_by_vanishing_poly(h))
    h_pieces = [h_coeffs[i:i + n] for i in range(0, len(h_coeffs) - n + 1, n)]  # par_chunks_exact(n)
    for piece in h_pieces:
        tr.write_point(params.commit(piece))
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)

    # ---- evaluations (:694-790)
    rot = domain.rotate_omega
    ev = o.eval_polynomial
    for instance_polys in inst_polys:
        for col, at in queries["Instance"]:
            tr.write_scalar(ev(instance_polys[col], rot(x, at)))
    for advice_polys in adv_polys:
        for col, at in queries["Advice"]:
            tr.write_scalar(ev(advice_polys[col], rot(x, at)))
    for col, at in queries["Fixed"]:
        tr.write_scalar(ev(pk.fixed_polys[col], rot(x, at)))
    h_poly = [0] * n                                                          # vanishing/prover.rs:119-123
    for piece in reversed(h_pieces):
        h_poly = [(a * xn + b) % R for a, b in zip(h_poly, piece)]
    tr.write_scalar(ev(random_poly, x))
    for poly in pk.sigma_polys:                                               # permutation/prover.rs:194-205
        tr.write_scalar(ev(poly, x))
    x_next, x_last = rot(x, 1), rot(x, -(bf + 1))
    for perm_polys in all_perm_polys:
        for i, z in enumerate(perm_polys):                                    # permutation/prover.rs:208-252
            tr.write_scalar(ev(z, x))
            tr.write_scalar(ev(z, x_next))
            if i + 1 < len(perm_polys):
                tr.write_scalar(ev(z, x_last))
    for lookups in all_lookups:
        for lk in lookups:                                                    # logup/prover.rs:421-447
            tr.write_scalar(ev(lk["m_poly"], x))
            for i, z in enumerate(lk["z_polys"]):
                tr.write_scalar(ev(z, x))
                tr.write_scalar(ev(z, x_next))
                if i + 1 < len(lk["z_polys"]):
                    tr.write_scalar(ev(z, x_last))
    for shuffle_polys in all_shuffle_polys:
        for z in shuffle_polys:                                               # shuffle/prover.rs:201-215
            tr.write_scalar(ev(z, x))
            tr.write_scalar(ev(z, x_next))

    # ---- queries for the multiopen argument (:792-838): (rotation, point, polynomial), circuit by circuit
    qs: List[Tuple[int, int, List[int]]] = []
    for ci in range(C):
        for col, at in queries["Instance"]:
            qs.append((at, rot(x, at), inst_polys[ci][col]))
        for col, at in queries["Advice"]:
            qs.append((at, rot(x, at), adv_polys[ci][col]))
        perm_polys = all_perm_polys[ci]
        for z in perm_polys:                                                  # permutation/prover.rs:257-303
            qs.append((0, x, z))
            qs.append((1, x_next, z))
        for z in list(reversed(perm_polys))[1:]:
            qs.append((-(bf + 1), x_last, z))
        for lk in all_lookups[ci]:                                            # logup/prover.rs:451-491
            qs.append((0, x, lk["m_poly"]))
            for z in lk["z_polys"]:
                qs.append((0, x, z))
                qs.append((1, x_next, z))
            for z in list(reversed(lk["z_polys"]))[1:]:
                qs.append((-(bf + 1), x_last, z))
        for z in all_shuffle_polys[ci]:                                       # shuffle/prover.rs:219-239
            qs.append((0, x, z))
            qs.append((1, x_next, z))
    for col, at in queries["Fixed"]:
        qs.append((at, rot(x, at), pk.fixed_polys[col]))
    for poly in pk.sigma_polys:                                               # permutation/prover.rs:182-192
        qs.append((0, x, poly))
    qs.append((0, x, h_poly))                                                 # vanishing/prover.rs:135-152
    qs.append((0, x, random_poly))

    if use_gwc:
        gwc_create_proof(params, tr, qs)
    else:
        shplonk_create_proof(params, tr, qs)
    return tr.finalize()


def construct_intermediate_sets(queries):
    """poly/multiopen/gwc.rs:38-62: BTreeMap<Rotation, Vec<Q>> -- queries grouped by ROTATION (this fork; upstream
    groups by point), groups in ascending rotation order, queries in arrival order, point = first query's."""
    groups: Dict[int, list] = {}
    for q in queries:
        groups.setdefault(q[0], []).append(q)
    return [(groups[r][0][1], groups[r]) for r in sorted(groups)]


def gwc_create_proof(params: Params, tr: Blake2bWrite, queries) -> None:
    """poly/multiopen/gwc/prover.rs:19-173 (the sort by batch size only schedules work; W's are written in set order)"""
    v = tr.squeeze_challenge()
    for z, group in construct_intermediate_sets(queries):
        poly_batch = [0] * params.n
        for q in group:
            assert q[1] == z
            poly_batch = [(a * v + b) % R for a, b in zip(poly_batch, q[2])]
        eval_batch = o.eval_polynomial(poly_batch, z)
        poly_batch[0] = (poly_batch[0] - eval_batch) % R
        witness = o.kate_division(poly_batch, z)
        tr.write_point(params.commit(witness))


# --------------------------------------------------------------------------
# SHPLONK (poly/multiopen/shplonk.rs, shplonk/prover.rs, shplonk/verifier.rs)
# --------------------------------------------------------------------------
def shplonk_intermediate_sets(queries, key_of, eval_of):
    """shplonk.rs:57-150.  queries: (rotation, point, commitment, ...); key_of(q): the identity of q's commitment
    (PolynomialPointer / CommitmentReference compare by address); eval_of(q_first_with_that_commitment, rotation).
    -> ([(commitments [(commitment, evals)], points)], super_point_set)"""
    rotation_point: Dict[int, int] = {}
    for q in queries:
        if rotation_point.setdefault(q[0], q[1]) != q[1]:
            raise AssertionError("rotation point matching consistency")            # :76
    super_point_set = [rotation_point[r] for r in sorted(rotation_point)]
    order: List[Tuple[object, tuple]] = []
    rots: Dict[object, set] = {}
    for q in queries:                                                             # :88-101
        k = key_of(q)
        if k not in rots:
            rots[k] = set()
            order.append((k, q))
        rots[k].add(q[0])
    groups: Dict[tuple, list] = {}
    for k, q in order:                                                            # :109-118 (BTreeMap<BTreeSet<Rotation>, _>)
        groups.setdefault(tuple(sorted(rots[k])), []).append((k, q))
    rotation_sets = []
    for rs in sorted(groups):                                                     # BTreeSet Ord = lexicographic
        commitments = [(q[2], [eval_of(k, r) for r in rs]) for k, q in groups[rs]]
        rotation_sets.append((commitments, [rotation_point[r] for r in rs]))
    return rotation_sets, super_point_set


def evaluate_vanishing_polynomial(roots: Sequence[int], z: int) -> int:
    """arithmetic.rs:905-923"""
    acc = 1
    for p in roots:
        acc = (z - p) * acc % R
    return acc


def _fold_polys(polys: Sequence[Sequence[int]], c: int, n: int) -> List[int]:
    acc = [0] * n
    for p in polys:
        acc = [(a * c + b) % R for a, b in zip(acc, p)]
    return acc


def shplonk_create_proof(params: Params, tr: Blake2bWrite, queries) -> None:
    """shplonk/prover.rs:78-234"""
    n = params.n
    y = tr.squeeze_challenge()
    polys = {id(q[2]): q[2] for q in queries}
    point_of = {q[0]: q[1] for q in queries}
    sets, super_point_set = shplonk_intermediate_sets(
        queries, lambda q: id(q[2]), lambda k, r: o.eval_polynomial(polys[k], point_of[r]))
    pad = lambda p: list(p) + [0] * (n - len(p))                                    # noqa: E731
    ext = [([(poly, pad(o.lagrange_interpolate(points, evals))) for poly, evals in commitments], points)
           for commitments, points in sets]                                       # :33-48 low_degree_equivalent
    v = tr.squeeze_challenge()
    quotients = []
    for commitments, points in ext:                                               # :97-128
        numerators = [[(a - b) % R for a, b in zip(poly, low)] for poly, low in commitments]
        n_x = _fold_polys(numerators, y, n)
        for p in points:                                                          # div_by_vanishing :20-26
            n_x = o.kate_division(n_x, p)
        quotients.append(pad(n_x))
    h_x = _fold_polys(quotients, v, n)
    tr.write_point(params.commit(h_x))
    u = tr.squeeze_challenge()
    zt_eval = evaluate_vanishing_polynomial(super_point_set, u)
    lins, z_diffs = [], []
    for commitments, points in ext:                                               # :163-192
        z_i = evaluate_vanishing_polynomial([p for p in super_point_set if p not in points], u)
        inner = []
        for poly, low in commitments:
            r_eval = o.eval_polynomial(low, u)
            inner.append([(poly[0] - r_eval) % R] + list(poly[1:]))
        l_i = _fold_polys(inner, y, n)
        lins.append([a * z_i % R for a in l_i])
        z_diffs.append(z_i)
    l_x = _fold_polys(lins, v, n)
    l_x = [(a - b * zt_eval) % R for a, b in zip(l_x, h_x)]
    assert o.eval_polynomial(l_x, u) == 0                                          # :213-216
    h2 = o.kate_division(l_x, u)
    inv = o.fr_inv(z_diffs[0])
    tr.write_point(params.commit([a * inv % R for a in h2]))


def shplonk_verify_proof(params, tr: Blake2bRead, queries) -> Tuple[Point, Point]:
    """shplonk/verifier.rs:23-104; queries: (rotation, point, commitment terms, eval, key)"""
    evals = {}
    for q in queries:
        evals.setdefault((q[4], q[0]), q[3])
    sets, super_point_set = shplonk_intermediate_sets(queries, lambda q: q[4], lambda k, r: evals[(k, r)])
    y = tr.squeeze_challenge()
    v = tr.squeeze_challenge()
    try:
        h1 = tr.read_point()
        u = tr.squeeze_challenge()
        h2 = tr.read_point()
    except TranscriptError:
        raise VerifyError("SamplingError")
    z_0_diff_inverse = z_0 = 0
    outer: List[List[Tuple[int, Point]]] = []
    r_outer_acc = 0
    for i, (commitments, points) in enumerate(sets):
        z_diff_i = evaluate_vanishing_polynomial([p for p in super_point_set if p not in points], u)
        if i == 0:
            z_0 = evaluate_vanishing_polynomial(points, u)
            z_0_diff_inverse = o.fr_inv(z_diff_i)
            z_diff_i = 1
        else:
            z_diff_i = z_diff_i * z_0_diff_inverse % R
        inner: List[Tuple[int, Point]] = []
        r_inner_acc = 0
        for terms, evs in commitments:
            r_x = o.lagrange_interpolate(points, evs)
            r_inner_acc = (y * r_inner_acc + o.eval_polynomial(r_x, u)) % R
            inner.append((1, _g1_lincomb(terms)))                                 # msm.eval() for the h commitment
        r_outer_acc = (v * r_outer_acc + r_inner_acc * z_diff_i) % R
        acc = 1
        scaled = []
        for s, p in reversed(inner):                                              # combine_with_base(y), msm.rs:136-145
            scaled.append((s * acc % R * z_diff_i % R, p))
            acc = acc * y % R
        outer.append(scaled)
    acc = 1
    right_terms: List[Tuple[int, Point]] = []
    for msm_terms in reversed(outer):                                             # combine_with_base(v), msm.rs:195-204
        right_terms += [(s * acc % R, p) for s, p in msm_terms]
        acc = acc * v % R
    right_terms += [((-r_outer_acc) % R, params.g1), ((-z_0) % R, h1), (u, h2)]
    return h2, _g1_lincomb(right_terms)


# --------------------------------------------------------------------------
# verify_proof
# --------------------------------------------------------------------------
class VerifyError(Exception):
    pass


def _g1_lincomb(terms: Sequence[Tuple[int, Point]]) -> Point:
    """evaluate an MSM<C> (poly/msm.rs): sum of scalar * point"""
    terms = [(s % R, p) for s, p in terms if p is not None and s % R]
    if not terms:
        return None
    return msm([s for s, _ in terms], o.g1_affine_encode([p for _, p in terms]))


def verify_proof(params: Params, vk: VerifyingKey, instances: Sequence[Sequence[int]], proof: bytes,
                 sign_bit: int = 7, pairing: bool = False, use_gwc: bool = True) -> bool:
    """one circuit instance per proof: verify_proof_multi with a list of one"""
    return verify_proof_multi(params, vk, [instances], proof, sign_bit, pairing, use_gwc)


def verify_proof_multi(params: Params, vk: VerifyingKey, instances: Sequence[Sequence[Sequence[int]]], proof: bytes,
                       sign_bit: int = 7, pairing: bool = False, use_gwc: bool = True) -> bool:
    """plonk/verifier.rs:127-507 with SingleVerifier, for `len(instances)` instances of the circuit in one proof
    (`instances: &[&[&[C::Scalar]]]`); False = the final check failed, VerifyError / TranscriptError = the proof is
    malformed.  `pairing=True` decides with the optimal-ate pairing on [s]G2 (Decider::verify); the default decides
    with the equivalent G1 equation [s]*left == right."""
    cs, domain = vk.cs, vk.domain
    n = params.n
    bf = cs.blinding_factors()
    queries = vk.queries
    C = len(instances)
    all_instance_commitments = []
    for circuit_instances in instances:
        if len(circuit_instances) != cs.num_instance:
            raise VerifyError("InvalidInstances")
        cm = []
        for inst in circuit_instances:                                        # :148-162
            if len(inst) > n - (bf + 1):
                raise VerifyError("InstanceTooLarge")
            cm.append(params.commit_lagrange([v % R for v in inst]))
        all_instance_commitments.append(cm)
    tr = Blake2bRead(proof, sign_bit)
    tr.common_scalar(vk.transcript_repr)                                      # :167
    for cm in all_instance_commitments:
        for c in cm:
            tr.common_point(c)
    all_advice_commitments = [[tr.read_point() for _ in range(cs.num_advice)] for _ in range(C)]
    theta = tr.squeeze_challenge()
    all_m_commitments = [[tr.read_point() for _ in cs.lookups] for _ in range(C)]
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    chunk_len = cs.degree() - 2
    n_sets = (len(cs.permutation_columns) + chunk_len - 1) // chunk_len
    all_perm_commitments = [[tr.read_point() for _ in range(n_sets)] for _ in range(C)]
    all_lookup_z_commitments = [[[tr.read_point() for _ in lk["input_expressions_sets"]] for lk in cs.lookups]
                                for _ in range(C)]
    all_shuffle_commitments = [[tr.read_point() for _ in cs.shuffles] for _ in range(C)]
    random_poly_commitment = tr.read_point()
    y = tr.squeeze_challenge()
    h_commitments = [tr.read_point() for _ in range(domain.quotient_poly_degree)]
    x = tr.squeeze_challenge()
    all_instance_evals = [[tr.read_scalar() for _ in queries["Instance"]] for _ in range(C)]
    all_advice_evals = [[tr.read_scalar() for _ in queries["Advice"]] for _ in range(C)]
    fixed_evals = [tr.read_scalar() for _ in queries["Fixed"]]
    random_eval = tr.read_scalar()
    permutation_evals = [tr.read_scalar() for _ in vk.permutation_commitments]
    all_perm_sets = []
    for perm_commitments in all_perm_commitments:
        perm_sets = []
        for i, c in enumerate(perm_commitments):                              # permutation/verifier.rs:74-101
            e, ne = tr.read_scalar(), tr.read_scalar()
            le = tr.read_scalar() if i + 1 < len(perm_commitments) else None
            perm_sets.append({"c": c, "eval": e, "next": ne, "last": le})
        all_perm_sets.append(perm_sets)
    all_lookups = []
    for m_commitments, lookup_z_commitments in zip(all_m_commitments, all_lookup_z_commitments):
        lookups = []
        for mc, zcs in zip(m_commitments, lookup_z_commitments):              # logup/verifier.rs:70-101
            m_eval = tr.read_scalar()
            zsets = []
            for i, c in enumerate(zcs):
                e, ne = tr.read_scalar(), tr.read_scalar()
                le = tr.read_scalar() if i + 1 < len(zcs) else None
                zsets.append({"c": c, "eval": e, "next": ne, "last": le})
            lookups.append({"m_c": mc, "m_eval": m_eval, "z": zsets})
        all_lookups.append(lookups)
    all_shuffles = [[{"c": c, "eval": tr.read_scalar(), "next": tr.read_scalar()} for c in shuffle_commitments]
                    for shuffle_commitments in all_shuffle_commitments]

    # ---- expected h(x) (:280-399)
    xn = pow(x, n, R)
    l_evals = domain.l_i_range(x, xn, range(-(bf + 1), 1))
    assert len(l_evals) == 2 + bf
    l_last = l_evals[0]
    l_blind = sum(l_evals[1:1 + bf]) % R
    l_0 = l_evals[1 + bf]
    active = (1 - (l_last + l_blind)) % R
    exprs: List[int] = []
    for ci in range(C):
        advice_evals, instance_evals = all_advice_evals[ci], all_instance_evals[ci]
        perm_sets, lookups, shuffles = all_perm_sets[ci], all_lookups[ci], all_shuffles[ci]
        ee = lambda e: eval_expression_at_queries(e, queries, fixed_evals, advice_evals, instance_evals)  # noqa: E731
        comp = lambda exprs: _fold([ee(e) for e in exprs], theta)             # noqa: E731

        def col_eval(column):
            kind, idx = column
            evals = {"Fixed": fixed_evals, "Advice": advice_evals, "Instance": instance_evals}[kind]
            return evals[queries[kind].index((idx, 0))]                      # get_any_query_index(column, cur)

        for gate in cs.gates:
            for poly in gate:
                exprs.append(ee(poly))
        # permutation/verifier.rs:105-203
        if perm_sets:
            exprs.append(l_0 * (1 - perm_sets[0]["eval"]) % R)
            last = perm_sets[-1]["eval"]
            exprs.append((last * last - last) * l_last % R)
            for i in range(1, len(perm_sets)):
                exprs.append((perm_sets[i]["eval"] - perm_sets[i - 1]["last"]) * l_0 % R)
            for si, st in enumerate(perm_sets):
                columns = cs.permutation_columns[si * chunk_len:(si + 1) * chunk_len]
                pevals = permutation_evals[si * chunk_len:(si + 1) * chunk_len]
                left = st["next"]
                for column, pe in zip(columns, pevals):
                    left = left * ((col_eval(column) + beta * pe + gamma) % R) % R
                right = st["eval"]
                cur_delta = beta * x % R * pow(P.FR_DELTA, si * chunk_len, R) % R
                for column in columns:
                    right = right * ((col_eval(column) + cur_delta + gamma) % R) % R
                    cur_delta = cur_delta * P.FR_DELTA % R
                exprs.append((left - right) * active % R)
        # logup/verifier.rs:104-218
        for lk, arg in zip(lookups, cs.lookups):
            zs = lk["z"]
            exprs.append(l_0 * zs[0]["eval"] % R)
            exprs.append(l_last * zs[-1]["eval"] % R)
            phi = [(comp(inp) + beta) % R for inp in arg["input_expressions_sets"][0]]
            tau = (comp(arg["table_expressions"]) + beta) % R
            product_fi = _prod(phi)
            sum_inv = sum(o.fr_inv(p) if p else 0 for p in phi) % R
            left = (tau * (zs[0]["next"] - zs[0]["eval"]) + lk["m_eval"]) % R * product_fi % R
            right = tau * product_fi % R * sum_inv % R
            exprs.append((left - right) * active % R)
            for i in range(1, len(zs)):
                exprs.append(l_0 * (zs[i]["eval"] - zs[i - 1]["last"]) % R)
            for zset, iset in list(zip(zs, arg["input_expressions_sets"]))[1:]:
                phi = [(comp(inp) + beta) % R for inp in iset]
                product_fi = _prod(phi)
                sum_inv = sum(o.fr_inv(p) if p else 0 for p in phi) % R
                exprs.append((zset["next"] - zset["eval"] - sum_inv) * product_fi % R * active % R)
        # shuffle/verifier.rs:58-121
        for sh, group in zip(shuffles, cs.shuffles):
            exprs.append(l_0 * (1 - sh["eval"]) % R)
            exprs.append(l_last * (sh["eval"] * sh["eval"] - sh["eval"]) % R)
            ps, pi = 1, 1
            for i, a in enumerate(group):
                ch = pow(beta, 1 + i, R)
                ps = ps * ((comp(a["shuffle_expressions"]) + ch) % R) % R
                pi = pi * ((comp(a["input_expressions"]) + ch) % R) % R
            exprs.append((sh["next"] * ps - sh["eval"] * pi) * active % R)
    expected_h_eval = _fold(exprs, y) * o.fr_inv((xn - 1) % R) % R          # vanishing/verifier.rs:88-96

    # h_commitment MSM: sum xn^i * h_i (vanishing/verifier.rs:98-106)
    h_terms, pw = [], 1
    for c in h_commitments:
        h_terms.append((pw, c))
        pw = pw * xn % R

    # ---- verifier queries (:401-488): (rotation, point, commitment terms, eval, identity of the commitment)
    rot = domain.rotate_omega
    x_next, x_last = rot(x, 1), rot(x, -(bf + 1))
    qs = []
    one = lambda c: [(1, c)]                                                   # noqa: E731
    for ci in range(C):
        perm_sets, lookups, shuffles = all_perm_sets[ci], all_lookups[ci], all_shuffles[ci]
        for (col, at), e in zip(queries["Instance"], all_instance_evals[ci]):
            qs.append((at, rot(x, at), one(all_instance_commitments[ci][col]), e, ("instance", ci, col)))
        for (col, at), e in zip(queries["Advice"], all_advice_evals[ci]):
            qs.append((at, rot(x, at), one(all_advice_commitments[ci][col]), e, ("advice", ci, col)))
        for i, st in enumerate(perm_sets):
            qs.append((0, x, one(st["c"]), st["eval"], ("perm", ci, i)))
            qs.append((1, x_next, one(st["c"]), st["next"], ("perm", ci, i)))
        for i, st in list(reversed(list(enumerate(perm_sets))))[1:]:
            qs.append((-(bf + 1), x_last, one(st["c"]), st["last"], ("perm", ci, i)))
        for li, lk in enumerate(lookups):
            qs.append((0, x, one(lk["m_c"]), lk["m_eval"], ("lookup_m", ci, li)))
            for i, st in enumerate(lk["z"]):
                qs.append((0, x, one(st["c"]), st["eval"], ("lookup_z", ci, li, i)))
                qs.append((1, x_next, one(st["c"]), st["next"], ("lookup_z", ci, li, i)))
            for i, st in list(reversed(list(enumerate(lk["z"]))))[1:]:
                qs.append((-(bf + 1), x_last, one(st["c"]), st["last"], ("lookup_z", ci, li, i)))
        for i, sh in enumerate(shuffles):
            qs.append((0, x, one(sh["c"]), sh["eval"], ("shuffle", ci, i)))
            qs.append((1, x_next, one(sh["c"]), sh["next"], ("shuffle", ci, i)))
    for (col, at), e in zip(queries["Fixed"], fixed_evals):
        qs.append((at, rot(x, at), one(vk.fixed_commitments[col]), e, ("fixed", col)))
    for i, (c, e) in enumerate(zip(vk.permutation_commitments, permutation_evals)):
        qs.append((0, x, one(c), e, ("sigma", i)))
    qs.append((0, x, h_terms, expected_h_eval, ("h",)))
    qs.append((0, x, one(random_poly_commitment), random_eval, ("random",)))

    if not use_gwc:
        left, right = shplonk_verify_proof(params, tr, qs)
        return Decider.verify(params, left, right) if pairing else Decider.verify_trapdoor(params, left, right)
    left, right = gwc_verify_proof(params, tr, qs)
    if pairing:
        return Decider.verify(params, left, right)
    return Decider.verify_trapdoor(params, left, right)


def _fold(vals: Sequence[int], c: int) -> int:
    acc = 0
    for v in vals:
        acc = (acc * c + v) % R
    return acc


def _prod(vals: Sequence[int]) -> int:
    acc = 1
    for v in vals:
        acc = acc * v % R
    return acc


def gwc_verify_proof(params: Params, tr: Blake2bRead, queries) -> Tuple[Point, Point]:
    """poly/multiopen/gwc/verifier.rs:16-91 -> (left, right) of the PairMSM, evaluated"""
    v = tr.squeeze_challenge()
    u = tr.squeeze_challenge()
    commitment_multi: List[Tuple[int, Point]] = []
    eval_multi = 0
    witness: List[Tuple[int, Point]] = []
    witness_with_aux: List[Tuple[int, Point]] = []
    scale = lambda terms, c: [(s * c % R, p) for s, p in terms]               # noqa: E731
    for z, group in construct_intermediate_sets(queries):
        try:
            wi = tr.read_point()
        except TranscriptError:
            raise VerifyError("SamplingError")
        witness_with_aux = scale(witness_with_aux, u) + [(z, wi)]
        witness = scale(witness, u) + [(1, wi)]
        commitment_multi = scale(commitment_multi, u)
        eval_multi = eval_multi * u % R
        commitment_batch: List[Tuple[int, Point]] = []
        eval_batch = 0
        for q in group:
            assert q[1] == z
            commitment_batch = scale(commitment_batch, v) + list(q[2])
            eval_batch = (eval_batch * v + q[3]) % R
        commitment_multi += commitment_batch
        eval_multi = (eval_multi + eval_batch) % R
    left = _g1_lincomb(witness)
    right = _g1_lincomb(witness_with_aux + commitment_multi + [(eval_multi, o.g1_neg(params.g1))])
    return left, right


class Decider:
    """poly/multiopen.rs:31-57: e(left, [s]G2) * e(right, -G2) == 1"""

    @staticmethod
    def verify_trapdoor(params: Params, left: Point, right: Point) -> bool:
        return o.g1_mul(left, params.s) == right

    @staticmethod
    def verify(params: Params, left: Point, right: Point) -> bool:
        from . import pairing as pr
        s_g2 = pr.g2_mul(pr.G2_GEN, params.s)                                # ParamsVerifier.s_g2 (commitment.rs:114-118)
        return pr.pairing_check([(left, s_g2), (right, pr.g2_neg(pr.G2_GEN))])
