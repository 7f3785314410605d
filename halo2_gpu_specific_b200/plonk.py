"""Host mirror of the prover that drives the engine: plonk::keygen (the parts the prover reads) and
plonk::create_proof / create_proof_from_witness of the reference, with every MSM, NTT, z column, quotient
evaluation, polynomial evaluation and multiopen fold issued as a batched call into the C ABI (SURVEY.md 8f
rank 2: "replace per-column par_iter calls with batched FFI calls; thread the caller's rng through").

  ConstraintSystem, queries        plonk/circuit.rs (the fields the prover reads; query_*_index / get_any_query_index)
  GraphBuilder -> Evaluator        plonk/evaluation.rs:307-448, 623-776 (Evaluator::new, add_expression)
  keygen                           plonk/keygen.rs:213-431, plonk/permutation/keygen.rs:196-262
  create_proof                     plonk/prover.rs:85-173 (instances), :916-1500 = :206-850 (the proof),
                                   plonk/vanishing/prover.rs:41-153, plonk/permutation/prover.rs:181-304,
                                   plonk/logup/prover.rs:70-256, 420-491, plonk/shuffle/prover.rs:200-240,
                                   poly/multiopen/gwc.rs:38-62, poly/multiopen/gwc/prover.rs:19-173
  create_proof_with_shplonk        plonk/prover.rs:1737-1757, poly/multiopen/shplonk.rs:57-150,
                                   poly/multiopen/shplonk/prover.rs:78-234
All paths relative to /root/reference/halo2_proofs/src.  Out of scope, as in DESIGN.md section 7: the circuit
front-end (layouter, selector compression, witness synthesis -- the advice columns arrive as fetch_witness would
deliver them) and the verifier.

Numbers are (.., 4) uint64 Montgomery arrays (the reference's in-memory Fr); challenges and evaluation points are
Python ints (canonical).  The numeric work goes through an engine object that speaks the block protocol described
above the engine classes: `ResidentEngine` (default: everything stays in HBM), `Engine` (every call copies its operands
through the host API) and the sharded variants of prover_sharded.py -- all of them device engines that raise without a
CUDA device; there is no CPU fallback.  The parameter also lets the host logic (transcript order, RNG order, query
grouping, multiplicities) be tested on CPU against a test double that lives in tests/.

Three inputs cannot be read from the reference tree and are parameters (see oracle/prover.py for the same list):
the verifying key's transcript scalar (`transcript_repr`), the point compression bit (`sign_bit`) and the source of
randomness (`rng`; draw order documented at create_proof).
"""
from __future__ import annotations

import hashlib
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error
from .transcript import Blake2bWrite, Point

R = _fr.R_MOD
DELTA = pow(_fr.GENERATOR, 1 << _fr.S, R)          # Fr::DELTA


# --------------------------------------------------------------------------
# ConstraintSystem (the part the prover reads)
# --------------------------------------------------------------------------
class ConstraintSystem:
    """Expressions are the reference's enum as tuples (see grand_product.ExprCompiler).  gates: list of
    polynomial lists (gate.polynomials()); lookups: [{"table_expressions", "input_expressions_sets"}]
    (logup::Argument); shuffles: groups of {"input_expressions", "shuffle_expressions"}; permutation_columns:
    [("Advice"|"Fixed"|"Instance", index)].  degree / blinding_factors are what cs.degree() /
    cs.blinding_factors() return (circuit.rs:1838-1944): computing them needs the front-end's query bookkeeping."""

    def __init__(self, num_fixed: int, num_advice: int, num_instance: int, degree: Optional[int] = None,
                 blinding_factors: Optional[int] = 5, gates=(), lookups=(), shuffles=(), permutation_columns=(),
                 advice_queries=None, fixed_queries=None, instance_queries=None, minimum_degree: Optional[int] = None):
        self.num_fixed, self.num_advice, self.num_instance = num_fixed, num_advice, num_instance
        self.gates = [list(g) for g in gates]
        self.lookups = list(lookups)
        self.shuffles = [list(g) for g in shuffles]
        self.permutation_columns = [tuple(c) for c in permutation_columns]
        self.minimum_degree = minimum_degree
        self._degree, self._blinding_factors = degree, blinding_factors
        self.advice_queries, self.fixed_queries, self.instance_queries = advice_queries, fixed_queries, instance_queries

    @staticmethod
    def expression_degree(e) -> int:
        """Expression::degree (circuit.rs)"""
        t = e[0]
        if t == "Constant":
            return 0
        if t in ("Fixed", "Advice", "Instance"):
            return 1
        if t in ("Negated", "Scaled"):
            return ConstraintSystem.expression_degree(e[1])
        if t == "Sum":
            return max(ConstraintSystem.expression_degree(e[1]), ConstraintSystem.expression_degree(e[2]))
        if t == "Product":
            return ConstraintSystem.expression_degree(e[1]) + ConstraintSystem.expression_degree(e[2])
        raise B2Error(B2_ERR_ARG, f"unknown Expression {e!r}")

    def compute_degree(self) -> int:
        """ConstraintSystem::degree (circuit.rs:1861-1915): the permutation argument needs 3
        (permutation.rs:29-62), a logup lookup max(4, 2 + input degree + table degree) (logup.rs:40-60), a shuffle
        2 + max(input, shuffle degree) (shuffle.rs:52-65), every gate polynomial its own degree; then minimum_degree"""
        deg = ConstraintSystem.expression_degree
        degree = 3
        for lk in self.lookups:
            inp = max([deg(e) for s in lk["input_expressions_sets"] for i in s for e in i] + [1])
            tab = max([deg(e) for e in lk["table_expressions"]] + [1])
            degree = max(degree, 4, 2 + inp + tab)
        for group in self.shuffles:
            for a in group:
                degree = max(degree, 2 + max([deg(e) for e in a["shuffle_expressions"]] + [1]),
                             2 + max([deg(e) for e in a["input_expressions"]] + [1]))
        for gate in self.gates:
            for poly in gate:
                degree = max(degree, deg(poly))
        return max(degree, self.minimum_degree or 1)

    def compute_blinding_factors(self) -> int:
        """ConstraintSystem::blinding_factors (circuit.rs:1919-1944): max(3, most queries of one advice column) + 2"""
        per_column: Dict[int, int] = {}
        for col, _ in self.queries()["Advice"]:
            per_column[col] = per_column.get(col, 0) + 1
        return max(3, max(per_column.values(), default=1)) + 2

    @classmethod
    def like(cls, other) -> "ConstraintSystem":
        """copy of any object with the same attributes (e.g. a front-end's description)"""
        return cls(other.num_fixed, other.num_advice, other.num_instance, other.degree(), other.blinding_factors(),
                   other.gates, other.lookups, other.shuffles, other.permutation_columns,
                   getattr(other, "advice_queries", None), getattr(other, "fixed_queries", None),
                   getattr(other, "instance_queries", None))

    def degree(self) -> int:
        if self._degree is None:
            self._degree = self.compute_degree()
        return self._degree

    def blinding_factors(self) -> int:
        if self._blinding_factors is None:
            self._blinding_factors = self.compute_blinding_factors()
        return self._blinding_factors

    def queries(self) -> Dict[str, List[Tuple[int, int]]]:
        """(column, rotation) lists per column kind.  Their order is the order of the circuit's meta.query_* calls
        in the reference; without a front-end the rule is: permutation columns at Rotation::cur (enable_equality),
        then gates, lookups (input sets, table), shuffles, in order of first appearance -- unless the lists were
        given explicitly."""
        if self.advice_queries is None:
            q: Dict[str, List[Tuple[int, int]]] = {"Advice": [], "Fixed": [], "Instance": []}

            def walk(e):
                t = e[0]
                if t in q:
                    if (e[1], e[2]) not in q[t]:
                        q[t].append((e[1], e[2]))
                elif t in ("Negated", "Scaled"):
                    walk(e[1])
                elif t in ("Sum", "Product"):
                    walk(e[1])
                    walk(e[2])
                elif t != "Constant":
                    raise B2Error(B2_ERR_ARG, f"unknown Expression {e!r}")

            for kind, col in self.permutation_columns:
                walk((kind, col, 0))
            for gate in self.gates:
                for poly in gate:
                    walk(poly)
            for lk in self.lookups:
                for s in lk["input_expressions_sets"]:
                    for inp in s:
                        for e in inp:
                            walk(e)
                for e in lk["table_expressions"]:
                    walk(e)
            for group in self.shuffles:
                for a in group:
                    for e in list(a["input_expressions"]) + list(a["shuffle_expressions"]):
                        walk(e)
            self.advice_queries, self.fixed_queries, self.instance_queries = q["Advice"], q["Fixed"], q["Instance"]
        return {"Advice": self.advice_queries, "Fixed": self.fixed_queries, "Instance": self.instance_queries}


# --------------------------------------------------------------------------
# Evaluator::new (evaluation.rs:307-448): expression DAG with common-subexpression sharing
# --------------------------------------------------------------------------
_RANK = {"Constant": 0, "Intermediate": 1, "Fixed": 2, "Advice": 3, "Instance": 4}
ZERO, ONE = ("Constant", 0), ("Constant", 1)


class GraphBuilder:
    """The interning tables of Evaluator: constants, rotations and calculations are appended on first use and
    referred to by index afterwards (evaluation.rs:623-668), so equal sub-expressions share one slot."""

    def __init__(self):
        self.constants: List[int] = []
        self.rotations: List[int] = []
        self.calculations: List[tuple] = []
        self._c: Dict[int, int] = {}
        self._r: Dict[int, int] = {}
        self._k: Dict[tuple, int] = {}

    def constant(self, v: int):
        v %= R
        if v not in self._c:
            self._c[v] = len(self.constants)
            self.constants.append(v)
        return ("Constant", self._c[v])

    def rotation(self, r: int) -> int:
        if r not in self._r:
            self._r[r] = len(self.rotations)
            self.rotations.append(r)
        return self._r[r]

    def calc(self, c: tuple):
        if c not in self._k:
            self._k[c] = len(self.calculations)
            self.calculations.append(c)
        return ("Intermediate", self._k[c])

    @staticmethod
    def _ordered(a, b):
        """commutative operands in derive(PartialOrd) order of ValueSource (evaluation.rs:45-58, 742, 758)"""
        return (a, b) if (_RANK[a[0]],) + tuple(a[1:]) <= (_RANK[b[0]],) + tuple(b[1:]) else (b, a)

    def expression(self, e):
        """add_expression, evaluation.rs:671-776 (including :727-728, which returns b for `0 - b`)"""
        t = e[0]
        if t == "Constant":
            return self.constant(e[1])
        if t in ("Fixed", "Advice", "Instance"):
            return self.calc(("Store", (t, e[1], self.rotation(e[2]))))
        if t == "Negated":
            if e[1][0] == "Constant":
                return self.constant(-e[1][1])
            a = self.expression(e[1])
            return a if a == ZERO else self.calc(("Negate", a))
        if t == "Sum":
            if e[2][0] == "Negated":
                a, b = self.expression(e[1]), self.expression(e[2][1])
                if a == ZERO:
                    return b
                return a if b == ZERO else self.calc(("Sub", a, b))
            a, b = self.expression(e[1]), self.expression(e[2])
            if a == ZERO:
                return b
            if b == ZERO:
                return a
            return self.calc(("Add",) + self._ordered(a, b))
        if t == "Product":
            a, b = self.expression(e[1]), self.expression(e[2])
            if a == ZERO or b == ZERO:
                return ZERO
            if a == ONE:
                return b
            if b == ONE:
                return a
            return self.calc(("Mul",) + self._ordered(a, b))
        if t == "Scaled":
            if e[2] % R == 0:
                return ZERO
            if e[2] % R == 1:
                return self.expression(e[1])
            c = self.constant(e[2])
            return self.calc(("Mul", self.expression(e[1]), c))
        raise B2Error(B2_ERR_ARG, f"unknown Expression {e!r}")

    def theta_fold(self, expressions):
        """evaluate_lc, evaluation.rs:350-360"""
        parts = [self.expression(x) for x in expressions]
        acc = parts[0]
        for p in parts[1:]:
            acc = self.calc(("LcTheta", acc, p))
        return acc


def evaluator_parts(cs) -> dict:
    """Evaluator::new (evaluation.rs:309-448, CPU structures; :539-578 for the shuffles) as plain data"""
    g = GraphBuilder()
    assert g.constant(0) == ZERO and g.constant(1) == ONE
    value_parts = [g.expression(poly) for gate in cs.gates for poly in gate]
    lookup_results = []
    for lk in cs.lookups:
        table = ("AddChallenge", g.theta_fold(lk["table_expressions"]), "Beta")
        sets = [[g.calc(("AddChallenge", g.theta_fold(inp), "Beta")) for inp in s] for s in lk["input_expressions_sets"]]
        products, sums = [], []
        for s in sets:
            acc = s[0]
            for v in s[1:]:
                acc = g.calc(("Mul", acc, v))
            products.append(("Store", acc))
        for s in sets:
            if len(s) == 1:
                sums.append(("Store", ONE))
                continue
            terms = []
            for i in range(len(s)):
                rest = s[:i] + s[i + 1:]
                acc = rest[0]
                for v in rest[1:]:
                    acc = g.calc(("Mul", acc, v))
                terms.append(acc)
            acc = terms[0]
            for v in terms[1:]:
                acc = g.calc(("Add", acc, v))
            sums.append(("Store", acc))
        lookup_results.append((table, products, sums))
    shuffle_results = []
    for group in cs.shuffles:
        lcs = [(g.theta_fold(a["input_expressions"]), g.theta_fold(a["shuffle_expressions"])) for a in group]
        pair = []
        for which in (0, 1):
            folded = [lc[which] for lc in lcs]
            acc = ("AddChallenge", folded[0], "Beta")
            for i, part in enumerate(folded[1:], start=1):
                acc = ("LcChallenge", part, g.calc(acc), "Beta", i + 1)
            pair.append(acc)
        shuffle_results.append(tuple(pair))
    return {"rotations": g.rotations, "constants": g.constants, "calculations": g.calculations,
            "value_parts": value_parts, "lookup_results": lookup_results, "shuffle_results": shuffle_results}


def build_evaluator(cs):
    from .evaluation import Evaluator
    p = evaluator_parts(cs)
    return Evaluator(p["rotations"], p["constants"], p["calculations"], p["value_parts"], p["lookup_results"],
                     p["shuffle_results"], cs.num_fixed, cs.num_advice, cs.num_instance, cs.permutation_columns,
                     cs.degree(), cs.blinding_factors())


# --------------------------------------------------------------------------
# randomness
# --------------------------------------------------------------------------
class SeededRng:
    """Reproducible source for the prover's random values (tests, benchmarks, the fixed-RNG parity the
    north star asks for).  fr_vec returns Montgomery limbs of 253-bit values (every 253-bit integer is below r)."""

    def __init__(self, seed: int):
        self._g = np.random.Generator(np.random.PCG64(seed))

    def u64_vec(self, n: int) -> np.ndarray:
        return self._g.integers(0, 1 << 64, size=n, dtype=np.uint64)

    def u16_vec(self, n: int) -> np.ndarray:
        return self._g.integers(0, 1 << 16, size=n, dtype=np.uint64)

    def fr_vec(self, n: int) -> np.ndarray:
        a = self._g.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)
        return a


class OsRng:
    """rand_core::OsRng for the same interface"""

    def u64_vec(self, n: int) -> np.ndarray:
        return np.frombuffer(os.urandom(8 * n), dtype=np.uint64).copy()

    def u16_vec(self, n: int) -> np.ndarray:
        return np.frombuffer(os.urandom(2 * n), dtype=np.uint16).astype(np.uint64)

    def fr_vec(self, n: int) -> np.ndarray:
        """uniform over Fr like Fr::random: 512 bits of OS entropy per element reduced mod r (bias < 2^-250)"""
        raw = os.urandom(64 * n)
        if n == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        return np.stack([_fr.to_mont(int.from_bytes(raw[64 * i:64 * i + 64], "little") % R) for i in range(n)])


class Blake2bRng:
    """Cryptographic stream for the same interface: BLAKE2b in counter mode keyed by a 256-bit seed (os.urandom by
    default).  Field elements are sampled by wide reduction (512 bits mod r, bias < 2^-250) and returned in Montgomery
    form, i.e. uniformly over Fr like the reference's Fr::random.  The multi-GPU prover gives every rank the SAME seed
    (prover_sharded.synchronized_rng: rank 0 draws it, broadcast), because every rank must blind identically."""

    def __init__(self, seed: Optional[bytes] = None):
        import hashlib
        self._seed = bytes(seed) if seed is not None else os.urandom(32)
        if len(self._seed) != 32:
            raise B2Error(B2_ERR_ARG, "Blake2bRng: the seed is 32 bytes")
        self._h, self._ctr = hashlib.blake2b, 0

    def _bytes(self, count: int) -> bytes:
        out = bytearray()
        while len(out) < count:
            out += self._h(self._ctr.to_bytes(8, "little"), key=self._seed, digest_size=64).digest()
            self._ctr += 1
        return bytes(out[:count])

    def u64_vec(self, n: int) -> np.ndarray:
        return np.frombuffer(self._bytes(8 * n), dtype=np.uint64).copy()

    def u16_vec(self, n: int) -> np.ndarray:
        return np.frombuffer(self._bytes(2 * n), dtype=np.uint16).astype(np.uint64)

    def fr_vec(self, n: int) -> np.ndarray:
        raw = self._bytes(64 * n)
        if n == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        return np.stack([_fr.to_mont(int.from_bytes(raw[64 * i:64 * i + 64], "little") % R) for i in range(n)])


# --------------------------------------------------------------------------
# engines
# --------------------------------------------------------------------------
# create_proof talks to its engine in terms of BLOCKS (a run of columns of n field elements that the engine
# owns) and COLUMNS (one column of a block).  What a block is belongs to the engine: the resident engine keeps
# device buffers, the host-API engine plain numpy arrays.  The operations:
#   put / alloc / cols / write_rows                      data movement
#   put_and_commit_lagrange, commit_lagrange, commit_lagrange_and_ifft, commit, lagrange_to_coeff
#   multiplicity_block                                   logup: compressed inputs / table -> the m(X) columns
#   permutation_z, logup_z, shuffle_z                    z columns written into columns of a block
#   random_poly, evaluate_h_blocks                       vanishing argument
#   eval_polynomial, poly_combine, sub_constant, kate_division_padded   evaluation phase and multiopen (GWC)
#   sub_low_degree, scale, sub_cols                      what SHPLONK adds
#   key_blocks, copy, stack, release, free


def chacha20_blocks(key: bytes, first: int, count: int) -> np.ndarray:
    """ChaCha20 key stream blocks first .. first + count - 1 (RFC 8439 block function, nonce 0) under a 32-byte key,
    as a (count, 16) array of little-endian 32-bit words; vectorised over the blocks.  The same function as
    csrc/chacha.cuh, which the device runs."""
    if len(key) != 32:
        raise B2Error(B2_ERR_ARG, "chacha20: the key is 32 bytes")
    kw = np.frombuffer(key, dtype="<u4")
    s = np.zeros((16, count), dtype=np.uint32)
    s[0], s[1], s[2], s[3] = 0x61707865, 0x3320646E, 0x79622D32, 0x6B206574
    for i in range(8):
        s[4 + i] = kw[i]
    s[12] = (np.arange(count, dtype=np.uint64) + np.uint64(first)).astype(np.uint32)
    x = s.copy()

    def rotl(v, c):
        return (v << np.uint32(c)) | (v >> np.uint32(32 - c))

    def qr(a, b, c, d):
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16)                             # noqa: E702
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12)                             # noqa: E702
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8)                              # noqa: E702
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7)                              # noqa: E702

    with np.errstate(over="ignore"):
        for _ in range(10):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)      # noqa: E702
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)      # noqa: E702
        x += s
    return np.ascontiguousarray(x.T)


def vanishing_streams(key: bytes, n: int):
    """The four per-coefficient streams of the vanishing argument's random polynomial from 256 bits of the caller's
    rng: ChaCha20 key stream under `key`, coefficient i takes blocks 3i, 3i + 1, 3i + 2; a_i = (block 3i as a 512-bit
    little-endian integer) mod r, b_i = (block 3i + 1) mod r, u_i / v_i = the first two 64-bit words of block 3i + 2.
    Returns (a_lo, a_hi, u, b_lo, b_hi, v): the 256-bit halves as (n, 4) little-endian 64-bit limbs (the reduction
    mod r is field arithmetic: a = a_lo + a_hi * 2^256), u and v as uint64.  Same generator as csrc/scan.cuh
    (vanishing_random_poly_kernel), which the device engines run instead."""
    w = chacha20_blocks(key, 0, 3 * n).reshape(n, 3, 16)
    limbs = lambda words: np.ascontiguousarray(words).view("<u8").astype(np.uint64).reshape(n, 4)     # noqa: E731
    third = np.ascontiguousarray(w[:, 2, :4]).view("<u8").astype(np.uint64).reshape(n, 2)
    return (limbs(w[:, 0, :8]), limbs(w[:, 0, 8:]), np.ascontiguousarray(third[:, 0]),
            limbs(w[:, 1, :8]), limbs(w[:, 1, 8:]), np.ascontiguousarray(third[:, 1]))


def canonical_max_bits(canonical: np.ndarray) -> int:
    """bit length of the largest value among canonical (n, 4) little-endian limbs"""
    c = np.asarray(canonical, dtype=np.uint64).reshape(-1, 4)
    for limb in (3, 2, 1, 0):
        top = int(c[:, limb].max()) if c.shape[0] else 0
        if top:
            return 64 * limb + top.bit_length()
    return 0


def _max_bits(counts: np.ndarray) -> int:
    """expression_max_bits, logup/prover.rs:493-517"""
    return max(16, int(counts.max()).bit_length() if counts.size else 0)


def _mont_vec(vals: Sequence[int]) -> np.ndarray:
    return np.stack([_fr.to_mont(int(v)) for v in vals]) if len(vals) else np.zeros((0, 4), np.uint64)


def _points(jac: np.ndarray) -> List[Point]:
    from .transcript import point_from_engine
    return [point_from_engine(p) for p in jac]


class ArrayBlocks:
    """Block protocol for engines whose blocks are numpy arrays (columns, n, 4): the host-API engine below, and the
    test double in tests/.  Everything here is expressed through the array primitives those engines provide
    (commit_lagrange, compress, permutation_commit, evaluate_h, kate_division, ...)."""

    def put(self, host: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(host, dtype=np.uint64)

    def alloc(self, count: int) -> np.ndarray:
        return np.zeros((count, self.domain.n, 4), dtype=np.uint64)

    @staticmethod
    def copy(block: np.ndarray) -> np.ndarray:
        return block.copy()

    @staticmethod
    def block_count(block: np.ndarray) -> int:
        return block.shape[0]

    @staticmethod
    def sub_block(block: np.ndarray, lo: int, hi: int) -> np.ndarray:
        return block[lo:hi]

    def commit_columns_with_bound(self, block: np.ndarray, max_bits: Optional[int]) -> List[Point]:
        """commit_lagrange_with_bound per column; max_bits None = find_max_scalar_bits per column first"""
        return self.put_and_commit_lagrange(block, max_bits)[1] if block.shape[0] else []

    @staticmethod
    def cols(block: np.ndarray) -> list:
        return [block[i] for i in range(block.shape[0])]

    @staticmethod
    def write_rows(col: np.ndarray, row: int, values: np.ndarray) -> None:
        col[row:row + len(values)] = values

    def put_and_commit_lagrange(self, host: np.ndarray, max_bits: Optional[int]):
        if max_bits is not None:
            return host, self.commit_lagrange(host, max_bits)
        points = []                                   # find_max_scalar_bits per column, plonk/prover.rs:945-962, 296
        for i in range(host.shape[0]):
            points += self.commit_lagrange(host[i:i + 1], canonical_max_bits(self.from_mont(host[i])))
        return host, points

    def multiplicity_block(self, cs, pk, advice, instance, theta: int, blinds, only=None):
        """logup `compress` (logup/prover.rs:70-256) for every lookup: compressed inputs and table on the engine,
        the multiplicities counted on the host as the reference does -> (block of m columns, bound for their commit).
        only: the lookup indices to work on (the multi-GPU prover divides them); the other columns are left zero."""
        n = self.domain.n
        usable = n - (cs.blinding_factors() + 1)
        m_canon = np.zeros((len(cs.lookups), n, 4), dtype=np.uint64)
        m_bits = 16
        for li, lk in enumerate(cs.lookups):
            if only is not None and li not in only:
                continue
            lists = [inp for s in lk["input_expressions_sets"] for inp in s] + [lk["table_expressions"]]
            comp = self.compress_canonical(lists, advice, pk.fixed_values, instance, theta)
            counts = logup_multiplicity(list(comp[:-1]), comp[-1], usable, n)
            m_bits = max(m_bits, _max_bits(counts[:usable]))
            m_canon[li, :, 0] = counts.astype(np.uint64)
            m_canon[li, usable:, 0] = blinds[li]
        return self.put_canonical(m_canon), m_bits

    def compress_canonical(self, expression_lists, advice, fixed, instance, theta: int) -> np.ndarray:
        comp = self.compress(expression_lists, advice, fixed, instance, theta)
        return self.from_mont(comp).reshape(len(expression_lists), self.domain.n, 4)

    def put_canonical(self, canonical: np.ndarray) -> np.ndarray:
        return self.to_mont(canonical).reshape(canonical.shape)

    def permutation_z(self, cs, pk, advice, instance, beta, gamma, blinds, out_cols) -> None:
        zs = self.permutation_commit(cs, pk.sigmas, advice, pk.fixed_values, instance, beta, gamma, blinds)
        for o, z in zip(out_cols, zs):
            o[:] = z

    def logup_z(self, cs, lookup, pk, advice, instance, m_col, theta, beta, out_cols) -> None:
        raw = self.logup_commit_z(cs, lookup, advice, pk.fixed_values, instance, m_col, theta, beta)
        for o, z in zip(out_cols, raw):
            o[:len(z)] = z

    def shuffle_z(self, cs, group, pk, advice, instance, theta, beta, out_col) -> None:
        z = self.shuffle_commit_product(cs, group, advice, pk.fixed_values, instance, theta, beta)
        out_col[:len(z)] = z

    def random_poly(self, random: np.ndarray, key: bytes) -> np.ndarray:
        """host restatement of vanishing_random_poly_kernel over the engine's element-wise field operations (the test
        double answers them with the oracle; the host-API device engine overrides this with the kernel itself)"""
        a_lo, a_hi, u, b_lo, b_hi, v = vanishing_streams(key, self.domain.n)
        kk = np.uint64(random.shape[0])
        # x mod r in Montgomery form: mont(R^2, lo) + mont(R^3, hi); the raw halves go second (fp_mul's bound)
        wide = lambda lo, hi: self.fr_vec("add", self.fr_vec("mul", np.broadcast_to(_RAW_R2, lo.shape), lo),      # noqa: E731
                                          self.fr_vec("mul", np.broadcast_to(_RAW_R3, hi.shape), hi))
        p = self.fr_vec("mul", self.fr_vec("add", wide(a_lo, a_hi), random[(u % kk).astype(np.int64)]),
                        self.fr_vec("add", wide(b_lo, b_hi), random[(v % kk).astype(np.int64)]))
        return p.reshape(1, -1, 4)

    def evaluate_h_blocks(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta) -> np.ndarray:
        n = self.domain.n
        lookups, pos = [], n_perm
        for li, cnt in enumerate(lookup_z_counts):
            lookups.append({"z": [z_block[pos + i] for i in range(cnt)], "m": m_block[li]})
            pos += cnt
        h = self.evaluate_h(pk, advice, instance, y, beta, gamma, theta, lookups,
                            [z_block[pos + i] for i in range(n_shuffles)], [z_block[i] for i in range(n_perm)])
        pieces = h.shape[0] // n                                          # par_chunks_exact(n)
        return np.ascontiguousarray(h[:pieces * n]).reshape(pieces, n, 4)

    @staticmethod
    def sub_constant(col: np.ndarray, value: int) -> None:
        col[0] = _fr.to_mont((_fr.from_mont(col[0]) - value) % R)

    def kate_division_padded(self, col: np.ndarray, z: int) -> np.ndarray:
        q = self.kate_division(col, z)
        return np.concatenate([q, np.zeros((1, 4), dtype=np.uint64)])

    @staticmethod
    def sub_low_degree(col: np.ndarray, coeffs: Sequence[int]) -> None:
        """col -= the polynomial with these (few) coefficients, in place"""
        for i, c in enumerate(coeffs):
            col[i] = _fr.to_mont((_fr.from_mont(col[i]) - c) % R)

    def scale(self, col: np.ndarray, s: int) -> np.ndarray:
        """s * col as a new column (the Horner fold of [col, 0] at s)"""
        return self.poly_combine([col, np.zeros_like(col)], s)

    def sub_cols(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        return self.fr_vec("sub", a, b).reshape(a.shape)

    @staticmethod
    def stack(cols) -> np.ndarray:
        return np.ascontiguousarray(np.stack(cols))

    @staticmethod
    def key_blocks(pk) -> dict:
        return {"fixed_values": pk.fixed_values, "fixed_polys": pk.fixed_polys, "sigmas": pk.sigmas,
                "sigma_polys": pk.sigma_polys}

    def release(self) -> None:
        pass

    def free(self) -> None:
        pass


class Engine(ArrayBlocks):
    """Host-API engine: every numeric step is one call into the C ABI with HOST arrays (each call copies its
    operands in and its result out), the way the reference's `cuda` build drives its GPU.  Used by keygen and
    available to create_proof (`engine=Engine(params, domain)`); create_proof's default is ResidentEngine."""

    def __init__(self, params, domain):
        from ._lib import require_gpu
        require_gpu()
        self.params, self.domain = params, domain

    _points = staticmethod(_points)

    def random_poly(self, random: np.ndarray, key: bytes) -> np.ndarray:
        """the device makes it (b2_vanishing_random_poly_dev), this engine's contract brings it back to the host"""
        import ctypes
        from ._lib import check, lib, ptr
        from .evaluation import DeviceBuffer
        n = self.domain.n
        random = np.ascontiguousarray(random, dtype=np.uint64).reshape(-1, 4)
        kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
        if kb.size != 32:
            raise B2Error(B2_ERR_ARG, "random_poly: the key is 32 bytes")
        buf = DeviceBuffer(n)
        try:
            check(lib().b2_vanishing_random_poly_dev(ptr(kb), ptr(random), random.shape[0], n, ctypes.c_void_p(buf.ptr), None))
            return buf.download().reshape(1, n, 4)
        finally:
            buf.free()

    # -- commitments
    def commit_lagrange(self, cols: np.ndarray, max_bits: int = _fr.NUM_BITS) -> List[Point]:
        """Params::commit_lagrange[_with_bound] per column (plonk/prover.rs:124-127, 293-299)"""
        return _points(self.params.commit_lagrange_batch(np.ascontiguousarray(cols), max_bits))

    def commit_lagrange_and_ifft(self, cols: np.ndarray) -> List[Point]:
        """Params::commit_lagrange_and_ifft per column (plonk/prover.rs:470-501, 535-553, 561-593); cols become
        coefficient forms in place"""
        d = self.domain
        return _points(self.params.commit_lagrange_batch(cols, ifft=(d.omega_inv, d.ifft_divisor)))

    def commit(self, cols: np.ndarray) -> List[Point]:
        """Params::commit per polynomial (vanishing/prover.rs:64, 86-96; gwc/prover.rs:162)"""
        return _points(self.params.commit_batch(cols))

    # -- transforms
    def lagrange_to_coeff(self, cols: np.ndarray) -> np.ndarray:
        return self.domain.lagrange_to_coeff_batch(cols)

    def coeff_to_extended(self, cols: np.ndarray) -> np.ndarray:
        return self.domain.coeff_to_extended(cols)

    def fft(self, a: np.ndarray) -> np.ndarray:
        """best_fft over the size-n domain, in place"""
        from .arithmetic import best_fft
        best_fft(a, self.domain.omega, self.domain.k)
        return a

    # -- element-wise field work
    def fr_vec(self, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        from ._lib import check, lib, ptr
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        if a.shape != b.shape:
            raise B2Error(B2_ERR_ARG, "fr_vec: shapes differ")
        out = np.empty_like(a)
        if a.shape[0]:
            check(lib().b2_field_vec(0, {"mul": 0, "add": 1, "sub": 2}[op], ptr(a), ptr(b), a.shape[0], ptr(out)))
        return out

    def to_mont(self, canonical: np.ndarray) -> np.ndarray:
        """canonical limbs -> Montgomery: the Montgomery product with R^2"""
        c = np.ascontiguousarray(canonical, dtype=np.uint64).reshape(-1, 4)
        return self.fr_vec("mul", c, np.broadcast_to(_RAW_R2, c.shape))

    def from_mont(self, a: np.ndarray) -> np.ndarray:
        """Montgomery -> canonical limbs: the Montgomery product with 1"""
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        return self.fr_vec("mul", a, np.broadcast_to(_RAW_ONE, a.shape))

    # -- z columns and the quotient
    def compress(self, expression_lists, advice, fixed, instance, theta: int) -> np.ndarray:
        from .grand_product import compress_expressions
        return compress_expressions(self.domain, expression_lists, advice, fixed, instance, theta)

    def permutation_commit(self, cs, sigmas, advice, fixed, instance, beta, gamma, blinds):
        from .grand_product import permutation_commit
        return permutation_commit(self.domain, cs.permutation_columns, cs.degree(), cs.blinding_factors(), sigmas, advice,
                                  fixed, instance, beta, gamma, blinds)

    def logup_commit_z(self, cs, lookup, advice, fixed, instance, m, theta, beta):
        from .grand_product import logup_commit_z
        return logup_commit_z(self.domain, lookup, cs.blinding_factors(), advice, fixed, instance, m, theta, beta)

    def shuffle_commit_product(self, cs, group, advice, fixed, instance, theta, beta):
        from .grand_product import shuffle_commit_product
        return shuffle_commit_product(self.domain, group, cs.blinding_factors(), advice, fixed, instance, theta, beta)

    def evaluate_h(self, pk, advice_polys, instance_polys, y, beta, gamma, theta, lookups, shuffles, permutations):
        """Evaluator::evaluate_h + divide_by_vanishing_poly + extended_to_coeff (plonk/prover.rs:663-690,
        vanishing/prover.rs:72-76) -> the n * (degree - 1) coefficients of h(X)"""
        return pk.ev.evaluate_h(self.domain, pk.fixed_polys, advice_polys, instance_polys, pk.l0, pk.l_last,
                                pk.l_active_row, pk.sigma_polys, y, beta, gamma, theta, lookups, shuffles, permutations,
                                to_coeff=True)

    # -- evaluation and opening
    def eval_polynomial(self, poly: np.ndarray, point: int) -> int:
        from .arithmetic import eval_polynomial
        return _fr.from_mont(eval_polynomial(poly, _fr.to_mont(point)))

    def poly_combine(self, polys, v: int) -> np.ndarray:
        from .arithmetic import poly_combine
        return poly_combine(polys, _fr.to_mont(v))

    def kate_division(self, poly: np.ndarray, z: int) -> np.ndarray:
        from .arithmetic import kate_division
        return kate_division(poly, _fr.to_mont(z))


_RAW_ONE = np.array([1, 0, 0, 0], dtype=np.uint64)                                     # canonical 1 (not Montgomery)
_RAW_R2 = np.array([(pow(2, 512, R) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
_RAW_R3 = np.array([(pow(2, 768, R) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class DevBlock:
    """`count` columns of n field elements in one device allocation (ResidentEngine); a block of one column is
    also what the engine hands out as a column"""
    __slots__ = ("ptr", "count", "n")

    def __init__(self, ptr: int, count: int, n: int):
        self.ptr, self.count, self.n = ptr, count, n

    def col(self, i: int) -> "DevBlock":
        return DevBlock(self.ptr + i * self.n * 32, 1, self.n)


class ResidentEngine:
    """Device-resident engine: every witness column crosses PCIe once (pipelined with its commitment), every other
    polynomial is born in HBM and stays there -- z columns, multiplicities, the random polynomial, h(X), the folded
    multiopen polynomials and their quotients; only commitments (96 B), evaluations (32 B) and one count per lookup
    (the bound for the commitment of m) go back -- the logup multiplicities are sorted and matched on the device too.
    The proving key's polynomials, and their evaluations on every coset of the extended domain, are made resident
    on first use and kept for later proofs (pk.fixed_cosets / permutation cosets are what the reference's CPU
    prover keeps too, plonk/keygen.rs)."""

    _PROFILED = ("put", "put_and_commit_lagrange", "commit_lagrange", "commit_lagrange_and_ifft", "commit",
                 "lagrange_to_coeff", "multiplicity_block", "permutation_z", "logup_z", "shuffle_z", "random_poly",
                 "evaluate_h_blocks", "eval_polynomial", "eval_polynomials", "poly_combine", "sub_constant", "kate_division_padded", "stack",
                 "sub_low_degree", "scale", "sub_cols", "key_blocks", "release")

    def __init__(self, params, domain, profile: bool = False):
        from ._lib import require_gpu
        require_gpu()
        self.params, self.domain = params, domain
        self.op_times: dict = {}       # profile=True: seconds and calls per engine operation
        if profile:
            import functools
            import time

            def timed(name, fn):
                @functools.wraps(fn)
                def run(*a, **kw):
                    t0 = time.perf_counter()
                    try:
                        return fn(*a, **kw)
                    finally:
                        rec = self.op_times.setdefault(name, [0.0, 0])
                        rec[0] += time.perf_counter() - t0
                        rec[1] += 1
                return run
            for name in self._PROFILED:
                setattr(self, name, timed(name, getattr(self, name)))
        from .evaluation import BufferPool
        self._pool = BufferPool()      # device blocks recycled between the steps of a proof and between proofs
        self._live: list = []          # per-proof device buffers, freed by release()
        self._kept: list = []          # proving-key data and constants, freed by free()
        self._keys: dict = {}          # id(pk) -> resident proving-key data (kept across proofs)
        self._consts: dict = {}
        self._early: dict = {}         # advice block ptr -> (coefficient forms, [coset evaluations]) made during the upload
        self._side = 0                 # side stream of the early transforms (created on first use)

    # -- memory
    def _buffer(self, elems: int, keep: bool = False):
        from .evaluation import DeviceBuffer
        b = DeviceBuffer(elems)
        (self._kept if keep else self._live).append(b)
        return b

    def alloc(self, count: int) -> DevBlock:
        n = self.domain.n
        return DevBlock(self._buffer(max(1, count) * n).ptr, count, n)

    def put(self, host: np.ndarray, keep: bool = False) -> DevBlock:
        host = np.ascontiguousarray(host, dtype=np.uint64)
        count, n = host.shape[0], host.shape[1]
        b = self._buffer(max(1, count) * n, keep)
        if count:
            b.upload(host)
        return DevBlock(b.ptr, count, n)

    def copy(self, block: DevBlock) -> DevBlock:
        import ctypes
        from ._lib import check, lib
        out = self.alloc(block.count)
        if block.count:
            check(lib().b2_memcpy_d2d(ctypes.c_void_p(out.ptr), ctypes.c_void_p(block.ptr), block.count * block.n * 32))
        return out

    @staticmethod
    def cols(block: DevBlock) -> list:
        return [block.col(i) for i in range(block.count)]

    @staticmethod
    def block_count(block: DevBlock) -> int:
        return block.count

    @staticmethod
    def sub_block(block: DevBlock, lo: int, hi: int) -> DevBlock:
        return DevBlock(block.ptr + lo * block.n * 32, hi - lo, block.n)

    def commit_columns_with_bound(self, block: DevBlock, max_bits: Optional[int]) -> List[Point]:
        """commit_lagrange_with_bound per resident column; max_bits None = the bound of each column is found on
        the device first (B2_MAX_BITS_AUTO)"""
        return self._commit(self.params.g_lagrange, 0, block, 0xFFFFFFFF if max_bits is None else max_bits, False)

    @staticmethod
    def write_rows(col: DevBlock, row: int, values: np.ndarray) -> None:
        from .grand_product import _h2d
        _h2d(col.ptr + row * 32, values)

    def get(self, block: DevBlock) -> np.ndarray:
        import ctypes
        from ._lib import check, lib
        out = np.empty((block.count, block.n, 4), dtype=np.uint64)
        check(lib().b2_memcpy_d2h(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(block.ptr), out.nbytes))
        return out

    def stack(self, cols) -> DevBlock:
        """columns -> one contiguous block (device-to-device copies)"""
        import ctypes
        from ._lib import check, lib
        n = self.domain.n
        out = self.alloc(len(cols))
        for i, c in enumerate(cols):
            check(lib().b2_memcpy_d2d(ctypes.c_void_p(out.ptr + i * n * 32), ctypes.c_void_p(c.ptr), n * 32))
        return out

    def _const_column(self, name: str, limbs: np.ndarray) -> int:
        """a resident column filled with one value (operand of the element-wise kernel)"""
        if name not in self._consts:
            n = self.domain.n
            self._consts[name] = self.put(np.broadcast_to(limbs, (1, n, 4)), keep=True)
        return self._consts[name].ptr

    def release(self) -> None:
        """free what the last proof allocated; the proving key stays resident"""
        self._drain_side()
        self._early = {}
        for b in self._live:
            b.free()
        self._live = []

    def free(self) -> None:
        """release() + the resident proving keys and constants"""
        self.release()
        for b in self._kept:
            b.free()
        self._kept = []
        self._keys, self._consts = {}, {}
        self._pool.trim()
        if self._side:
            from ._lib import lib
            lib().b2_stream_destroy(self._side)
            self._side = 0

    # -- commitments
    def _commit(self, srs, host_ptr, block: DevBlock, max_bits: int, ifft: bool) -> List[Point]:
        import ctypes
        from ._lib import check, lib, ptr
        d = self.domain
        out = np.zeros((block.count, 12), dtype=np.uint64)
        if block.count:
            vp = ctypes.c_void_p
            check(lib().b2_commit_batch_resident(srs.handle, vp(host_ptr) if host_ptr else None, 0 if host_ptr else 1,
                                                 vp(block.ptr), block.count, block.n, int(max_bits), 1 if ifft else 0,
                                                 ptr(d.omega_inv), ptr(d.ifft_divisor), d.k, ptr(out)))
        return _points(out)

    def put_and_commit_lagrange(self, host: np.ndarray, max_bits: Optional[int]):
        """host columns -> resident block, committed on the way in (copy of column i + 1 overlaps the MSM of column i)"""
        if not (host.flags.c_contiguous and host.dtype == np.uint64 and host.ndim == 3):
            raise B2Error(B2_ERR_ARG, "expected a C-contiguous uint64 (columns, n, 4) array")
        # no bound given: find_max_scalar_bits per column on the device (plonk/prover.rs:945-962, 296), inside the
        # same pipelined call (B2_MAX_BITS_AUTO): scan of column c after its copy, its MSM under the copy of c + 1
        bits = 0xFFFFFFFF if max_bits is None else max_bits
        count, n = host.shape[0], host.shape[1]
        block = self.alloc(count)
        nc = 1 << (self.domain.extended_k - self.domain.k)
        if not self.EARLY_TRANSFORMS or count < 8 or (nc + 1) * count * n * 32 > self.EARLY_TRANSFORM_BYTES:
            return block, self._commit(self.params.g_lagrange, host.ctypes.data, block, bits, False)
        # The upload is bound by PCIe and leaves the multiplier pipe idle about half of the time, while evaluate_h will
        # later need every advice polynomial on every coset of the extended domain -- transforms that depend on the
        # witness only, not on any challenge.  So the columns go up in groups, and while group g + 1 crosses PCIe the
        # side stream turns group g into coefficient form and into its coset evaluations (lagrange_to_coeff and
        # evaluate_h_blocks pick them up from self._early).
        return block, self.put_columns_with_early_transforms(block, host, 0, count, bits, range(nc))

    def put_columns_with_early_transforms(self, block: DevBlock, host: np.ndarray, lo: int, hi: int, bits: int,
                                          coset_ids) -> List[Point]:
        """columns [lo, hi) of `host` -> the same columns of `block`, committed on the way in, in groups; behind every
        group the side stream makes its coefficient forms and its evaluations on the cosets `coset_ids`.  Registers the
        results under the block: lagrange_to_coeff and evaluate_h_blocks use what is there and compute the rest."""
        count, n = block.count, block.n
        coeff = self.alloc(count)
        cosets = {c: self.alloc(count) for c in coset_ids}
        step = max(2, (hi - lo + 7) // 8)
        points: List[Point] = []
        for g_lo in range(lo, hi, step):
            g_hi = min(hi, g_lo + step)
            points += self._commit(self.params.g_lagrange, host.ctypes.data + g_lo * n * 32,
                                   self.sub_block(block, g_lo, g_hi), bits, False)
            self._early_transforms(block, coeff, cosets, g_lo, g_hi)
        self._early[block.ptr] = {"coeff": coeff, "lo": lo, "hi": hi, "cosets": cosets}
        return points

    EARLY_TRANSFORMS = True                  # class switch (the sharded engines divide the columns differently)
    EARLY_TRANSFORM_BYTES = 64 << 30         # HBM the early coefficient forms + coset evaluations may take

    def _side_stream(self) -> int:
        if not self._side:
            import ctypes
            from ._lib import check, lib
            st = ctypes.c_void_p()
            check(lib().b2_stream_create(ctypes.byref(st)))
            self._side = st.value
        return self._side

    def _drain_side(self) -> None:
        if self._side:
            from ._lib import check, lib
            check(lib().b2_stream_synchronize(self._side))

    def _early_transforms(self, block: DevBlock, coeff: DevBlock, cosets, lo: int, hi: int) -> None:
        """columns [lo, hi) of a Lagrange block -> coefficient forms and coset evaluations, asynchronously"""
        import ctypes
        from ._lib import NttDesc, check, lib
        from .evaluation import coeff_to_coset_dev
        dm, st = self.domain, self._side_stream()
        off = lo * dm.n * 32
        d = NttDesc()
        d.log_n, d.location = dm.k, 1
        d.omega, d.divisor = dm.omega_inv.ctypes.data, dm.ifft_divisor.ctypes.data
        d.n_in = d.n_out = d.in_stride = d.out_stride = dm.n
        d.columns = hi - lo
        d.in_, d.out = block.ptr + off, coeff.ptr + off
        d.stream = st
        check(lib().b2_ntt_exec(ctypes.byref(d)))
        for c, cos in cosets.items():
            g_c = dm._zeta * pow(dm._ext_omega, c, R) % R
            coeff_to_coset_dev(dm, coeff.ptr + off, hi - lo, g_c, cos.ptr + off, stream=st)

    def commit_lagrange(self, block: DevBlock, max_bits: int = _fr.NUM_BITS) -> List[Point]:
        return self._commit(self.params.g_lagrange, 0, block, max_bits, False)

    def commit_lagrange_and_ifft(self, block: DevBlock) -> List[Point]:
        return self._commit(self.params.g_lagrange, 0, block, _fr.NUM_BITS, True)

    def commit(self, block: DevBlock) -> List[Point]:
        return self._commit(self.params.g, 0, block, _fr.NUM_BITS, False)

    # -- transforms
    def lagrange_to_coeff(self, block: DevBlock) -> DevBlock:
        import ctypes
        from ._lib import NttDesc, check, lib
        cb = block.n * 32
        early, e_lo, e_hi = None, 0, 0
        for base, e in self._early.items():
            # `block` may be the registered block itself or a run of its columns (the sharded engines pass their share)
            if base <= block.ptr < base + e["coeff"].count * cb and (block.ptr - base) % cb == 0:
                first = (block.ptr - base) // cb
                e_lo, e_hi = max(e["lo"], first) - first, min(e["hi"], first + block.count) - first
                if e_hi > e_lo:
                    early = (e, first)
                break
        runs = [(0, block.count)]
        if early is not None:
            # these coefficient forms were made while the columns were uploaded (put_columns_with_early_transforms)
            e, first = early
            self._drain_side()
            check(lib().b2_memcpy_d2d(ctypes.c_void_p(block.ptr + e_lo * cb),
                                      ctypes.c_void_p(e["coeff"].ptr + (first + e_lo) * cb), (e_hi - e_lo) * cb))
            runs = [(0, e_lo), (e_hi, block.count)]
        dm = self.domain
        for lo, hi in runs:
            if hi <= lo:
                continue
            d = NttDesc()
            d.log_n, d.location = dm.k, 1
            d.omega, d.divisor = dm.omega_inv.ctypes.data, dm.ifft_divisor.ctypes.data
            d.n_in = d.n_out = d.in_stride = d.out_stride = dm.n
            d.columns = hi - lo
            d.in_ = d.out = block.ptr + lo * cb
            check(lib().b2_ntt_exec(ctypes.byref(d)))
        return block

    # -- element-wise
    def _fr_vec(self, op: int, a_ptr: int, b_ptr: int, count: int, out_ptr: int) -> None:
        import ctypes
        from ._lib import check, lib
        vp = ctypes.c_void_p
        check(lib().b2_fr_vec_dev(op, vp(a_ptr), vp(b_ptr), count, vp(out_ptr), None))

    # -- logup
    def _columns(self, pk, advice: DevBlock, instance: DevBlock):
        from .grand_product import _Columns
        key = self.key_blocks(pk)
        ptrs = lambda b: [b.ptr + i * b.n * 32 for i in range(b.count)]      # noqa: E731
        return _Columns.resident(ptrs(key["fixed_values"]), ptrs(advice), ptrs(instance), self.domain.n)

    def multiplicity_block(self, cs, pk, advice, instance, theta: int, blinds, only=None):
        """logup `compress` (logup/prover.rs:70-256) for every lookup without leaving the device: the compressed
        inputs and table are produced by the expression kernel, sorted and matched there (logup_multiplicity_device),
        and m(X) is written as a resident column; only the largest count (for the commit bound) returns.
        only: the lookup indices to work on (the multi-GPU prover divides them); other columns are left untouched."""
        from .grand_product import compress_expressions_dev
        n = self.domain.n
        bf = cs.blinding_factors()
        usable = n - (bf + 1)
        n_lookups = len(cs.lookups)
        ms = self.alloc(n_lookups)
        m_bits = 16
        cols = self._columns(pk, advice, instance)
        for li, lk in enumerate(cs.lookups):
            if only is not None and li not in only:
                continue
            lists = [inp for s in lk["input_expressions_sets"] for inp in s] + [lk["table_expressions"]]
            comp = self.alloc(len(lists))
            compress_expressions_dev(self.domain, lists, cols, theta, comp.ptr)
            m_col = ms.col(li)
            largest = logup_multiplicity_device(comp.ptr, len(lists) - 1, comp.ptr + (len(lists) - 1) * n * 32, usable, n,
                                                m_col.ptr)
            m_bits = max(m_bits, largest.bit_length())
            self.write_rows(m_col, usable, _mont_vec(blinds[li]))              # logup/prover.rs:232-236
        return ms, m_bits

    # -- z columns
    def permutation_z(self, cs, pk, advice, instance, beta, gamma, blinds, out_cols) -> None:
        from .grand_product import permutation_commit_dev
        sig = self.key_blocks(pk)["sigmas"]
        permutation_commit_dev(self.domain, cs.permutation_columns, cs.degree(), cs.blinding_factors(),
                               [sig.ptr + j * sig.n * 32 for j in range(sig.count)], self._columns(pk, advice, instance),
                               beta, gamma, blinds, [c.ptr for c in out_cols])

    def logup_z(self, cs, lookup, pk, advice, instance, m_col, theta, beta, out_cols) -> None:
        from .grand_product import logup_commit_z_dev
        logup_commit_z_dev(self.domain, lookup, cs.blinding_factors(), self._columns(pk, advice, instance), m_col.ptr,
                           theta, beta, [c.ptr for c in out_cols])

    def shuffle_z(self, cs, group, pk, advice, instance, theta, beta, out_col) -> None:
        from .grand_product import shuffle_commit_product_dev
        shuffle_commit_product_dev(self.domain, group, cs.blinding_factors(), self._columns(pk, advice, instance), theta,
                                   beta, out_col.ptr)

    # -- vanishing argument
    def random_poly(self, random: np.ndarray, key: bytes) -> DevBlock:
        """generated where it is used (b2_vanishing_random_poly_dev): the polynomial never exists on the host"""
        import ctypes
        from ._lib import check, lib, ptr
        out = self.alloc(1)
        random = np.ascontiguousarray(random, dtype=np.uint64).reshape(-1, 4)
        kb = np.frombuffer(bytes(key), dtype=np.uint8).copy()
        if kb.size != 32:
            raise B2Error(B2_ERR_ARG, "random_poly: the key is 32 bytes")
        check(lib().b2_vanishing_random_poly_dev(ptr(kb), ptr(random), random.shape[0], self.domain.n,
                                                 ctypes.c_void_p(out.ptr), None))
        return out

    def key_blocks(self, pk) -> dict:
        """the proving key's columns, resident (made so on first use)"""
        key = self._keys.get(id(pk))
        if key is None:
            key = {name: self.put(getattr(pk, name), keep=True)
                   for name in ("fixed_values", "fixed_polys", "sigmas", "sigma_polys")}
            key["pk"] = pk          # keeps the key object alive, so that its id cannot be reused while cached
            self._keys[id(pk)] = key
        return key

    def _key_cosets(self, pk) -> list:
        """per coset c of the extended domain: one block holding the evaluations of the fixed and sigma polynomials
        on it, followed by the rows of l0 / l_last / l_active_row that belong to it"""
        from .evaluation import coeff_to_coset_dev
        key = self.key_blocks(pk)
        if "cosets" not in key:
            dm = self.domain
            n = dm.n
            nc = 1 << (dm.extended_k - dm.k)
            F, S = key["fixed_polys"].count, key["sigma_polys"].count
            out = []
            for c in range(nc):
                g_c = dm._zeta * pow(dm._ext_omega, c, R) % R
                blk = DevBlock(self._buffer((F + S + 3) * n, keep=True).ptr, F + S + 3, n)
                if F:
                    coeff_to_coset_dev(dm, key["fixed_polys"].ptr, F, g_c, blk.ptr)
                if S:
                    coeff_to_coset_dev(dm, key["sigma_polys"].ptr, S, g_c, blk.ptr + F * n * 32)
                for i, v in enumerate((pk.l0, pk.l_last, pk.l_active_row)):
                    rows = np.ascontiguousarray(np.asarray(v, dtype=np.uint64).reshape(-1, 4)[c::nc])
                    self.write_rows(blk.col(F + S + i), 0, rows)
                out.append(blk)
            key["cosets"] = out
        return key["cosets"]

    def evaluate_h_blocks(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta, tasks=None, combine=None) -> DevBlock:
        """Evaluator::evaluate_h (plonk/evaluation.rs:778-1226) coset by coset from resident coefficient forms,
        divide_by_vanishing_poly folded into the store, extended_to_coeff on the device (vanishing/prover.rs:72-76):
        -> the h(X) pieces as a resident block.

        tasks: None = the whole extended domain; else (coset, row_begin, row_count) triples (parallel.quotient_tasks):
        only those rows are evaluated, into a zeroed buffer, and `combine(hext_block)` must complete it before the
        inverse transform (the multi-GPU split: an all-reduce of disjoint supports, prover_sharded.py)."""
        import ctypes
        from ._lib import NttDesc, check, lib
        from .evaluation import coeff_to_coset_dev
        dm = self.domain
        n = dm.n
        nc = 1 << (dm.extended_k - dm.k)
        if tasks is None:
            tasks = [(c, 0, n) for c in range(nc)]
            partial = False
        else:
            partial = True
        key_cosets = self._key_cosets(pk)
        F, S = pk.fixed_polys.shape[0], pk.sigma_polys.shape[0]
        prog = pk.ev.program(n_perm, list(lookup_z_counts), n_shuffles)
        challenges = [beta % R, gamma % R, theta % R, y % R]
        d = beta * dm._zeta % R                                               # delta_start, evaluation.rs:1011
        for _ in range(S):
            challenges.append(d)
            d = d * DELTA % R
        early = self._early.get(advice.ptr)        # the advice polynomials' coset evaluations may exist already
        if early is not None:
            self._drain_side()
        witness = [b for b in (advice, instance, z_block, m_block) if b.count]
        # (the advice block's coset evaluations live in the early buffers where those exist for the coset at hand)
        cos = {id(b): self.alloc(b.count) for b in witness
               if not (b is advice and early and all(c in early["cosets"] for c, _, _ in tasks))}
        hext = self._buffer(dm.extended_len())
        if partial:
            self._fr_vec(2, hext.ptr, hext.ptr, dm.extended_len(), hext.ptr)          # x - x: a zeroed buffer
        ptrs = lambda b: [cos[id(b)].ptr + i * n * 32 for i in range(b.count)] if b.count else []     # noqa: E731
        current, adv_cos = None, None
        for c, row_begin, row_count in tasks:
            if c != current:                           # tasks of one coset are adjacent (coset-major order)
                g_c = dm._zeta * pow(dm._ext_omega, c, R) % R
                for b in witness:
                    if b is advice and early and c in early["cosets"]:
                        # columns [lo, hi) of this coset were evaluated during the upload; the others (a sharded
                        # engine's peers uploaded them) are computed now, into the same buffer
                        adv_cos = early["cosets"][c]
                        for r_lo, r_hi in ((0, early["lo"]), (early["hi"], b.count)):
                            if r_hi > r_lo:
                                coeff_to_coset_dev(dm, b.ptr + r_lo * n * 32, r_hi - r_lo, g_c, adv_cos.ptr + r_lo * n * 32)
                        continue
                    if b is advice:
                        adv_cos = cos[id(b)]
                    coeff_to_coset_dev(dm, b.ptr, b.count, g_c, cos[id(b)].ptr)
                current = c
            kc = key_cosets[c]
            kp = [kc.ptr + i * n * 32 for i in range(kc.count)]
            zp, mp = ptrs(z_block), ptrs(m_block)
            ap = [adv_cos.ptr + i * n * 32 for i in range(advice.count)] if advice.count else []
            aux = kp[F + S:F + S + 3] + kp[F:F + S] + zp[:n_perm]
            pos = n_perm
            for li, cnt in enumerate(lookup_z_counts):
                aux += zp[pos:pos + cnt] + [mp[li]]
                pos += cnt
            aux += zp[pos:pos + n_shuffles]
            full = row_begin == 0 and row_count == n
            prog.eval(dm.k, 1, kp[:F], ap, ptrs(instance), aux, challenges, hext.ptr,
                      x0=pow(dm._ext_omega, c, R), x_step=dm._omega, scale=dm.t_evaluations[c:c + 1],
                      out_stride=nc, out_offset=c if full else c + row_begin * nc,   # row i lands at i * nc + c
                      row_begin=0 if full else row_begin, row_count=0 if full else row_count)
        if combine is not None:
            combine(DevBlock(hext.ptr, 1, dm.extended_len()))
        pieces = dm.quotient_poly_degree
        hcoef = self.alloc(pieces)
        z = np.concatenate([dm.g_coset_inv, dm.g_coset])                      # leaving the coset: {zeta^2, zeta}
        t = NttDesc()
        t.log_n, t.location = dm.extended_k, 1
        t.omega, t.divisor = dm.extended_omega_inv.ctypes.data, dm.extended_ifft_divisor.ctypes.data
        t.coset_out = z.ctypes.data
        t.n_in = t.in_stride = dm.extended_len()
        t.n_out = t.out_stride = n * pieces
        t.columns, t.in_, t.out = 1, hext.ptr, hcoef.ptr
        check(lib().b2_ntt_exec(ctypes.byref(t)))
        return hcoef

    # -- evaluation and opening
    def eval_polynomial(self, col: DevBlock, point: int) -> int:
        import ctypes
        from ._lib import check, lib, ptr
        out = np.empty(4, dtype=np.uint64)
        pt = _fr.to_mont(point)
        check(lib().b2_eval_polynomial_dev(ctypes.c_void_p(col.ptr), 1, col.n, col.n, ptr(pt), ptr(out)))
        return _fr.from_mont(out)

    def eval_polynomials(self, cols, point: int) -> List[int]:
        """every polynomial of `cols` at one point with one launch (b2_eval_polynomials_dev)"""
        import ctypes
        from ._lib import check, lib, ptr
        if not cols:
            return []
        out = np.empty((len(cols), 4), dtype=np.uint64)
        pt = _fr.to_mont(point)
        arr = (ctypes.c_void_p * len(cols))(*[c.ptr for c in cols])
        check(lib().b2_eval_polynomials_dev(arr, len(cols), cols[0].n, ptr(pt), ptr(out)))
        return [_fr.from_mont(out[i]) for i in range(len(cols))]

    def poly_combine(self, cols, v: int) -> DevBlock:
        import ctypes
        from ._lib import check, lib, ptr
        out = self.alloc(1)
        arr = (ctypes.c_void_p * len(cols))(*[c.ptr for c in cols])
        vm = _fr.to_mont(v)
        check(lib().b2_poly_combine_dev(arr, len(cols), self.domain.n, ptr(vm), ctypes.c_void_p(out.ptr), None))
        return out

    def sub_constant(self, col: DevBlock, value: int) -> None:
        import ctypes
        from ._lib import check, lib
        c0 = np.empty(4, dtype=np.uint64)
        check(lib().b2_memcpy_d2h(ctypes.c_void_p(c0.ctypes.data), ctypes.c_void_p(col.ptr), 32))
        self.write_rows(col, 0, _fr.to_mont((_fr.from_mont(c0) - value) % R).reshape(1, 4))

    def sub_low_degree(self, col: DevBlock, coeffs: Sequence[int]) -> None:
        """col -= the polynomial with these (few) coefficients, in place: a few dozen bytes each way"""
        import ctypes
        from ._lib import check, lib
        m = len(coeffs)
        if m == 0:
            return
        head = np.empty((m, 4), dtype=np.uint64)
        check(lib().b2_memcpy_d2h(ctypes.c_void_p(head.ctypes.data), ctypes.c_void_p(col.ptr), m * 32))
        self.write_rows(col, 0, _mont_vec([(_fr.from_mont(head[i]) - c) % R for i, c in enumerate(coeffs)]))

    def scale(self, col: DevBlock, s: int) -> DevBlock:
        """s * col as a new column (the Horner fold of [col, 0] at s)"""
        zero = DevBlock(self._const_column("zero", np.zeros(4, dtype=np.uint64)), 1, self.domain.n)
        return self.poly_combine([col, zero], s)

    def sub_cols(self, a: DevBlock, b: DevBlock) -> DevBlock:
        out = self.alloc(1)
        self._fr_vec(2, a.ptr, b.ptr, self.domain.n, out.ptr)
        return out

    def kate_division_padded(self, col: DevBlock, z: int) -> DevBlock:
        """kate_division (n - 1 coefficients) into a column of n with a zero on top, so that it commits like one"""
        import ctypes
        from ._lib import check, lib, ptr
        n = self.domain.n
        out = self.alloc(1)
        zm = _fr.to_mont(z)
        check(lib().b2_kate_division_dev(ctypes.c_void_p(col.ptr), n, ptr(zm), ctypes.c_void_p(out.ptr), None))
        self.write_rows(out, n - 1, np.zeros((1, 4), dtype=np.uint64))
        return out


# --------------------------------------------------------------------------
# keygen
# --------------------------------------------------------------------------
class VerifyingKey:
    def __init__(self, cs, domain, fixed_commitments, permutation_commitments, transcript_repr: Optional[int]):
        self.cs, self.domain = cs, domain
        self.fixed_commitments, self.permutation_commitments = fixed_commitments, permutation_commitments
        if transcript_repr is None:
            # plonk.rs:91-109 hashes Rust's Debug text of the pinned key; that text cannot be produced here, so the
            # default is the same construction over this package's own description of the key
            s = repr((domain.k, domain.extended_k, hex(domain._omega), fixed_commitments, permutation_commitments,
                      cs.num_fixed, cs.num_advice, cs.num_instance, cs.gates, cs.lookups, cs.shuffles,
                      cs.permutation_columns, cs.queries(), cs.degree(), cs.blinding_factors())).encode()
            h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
            h.update(len(s).to_bytes(8, "little"))
            h.update(s)
            transcript_repr = int.from_bytes(h.digest(), "little")
            global _WARNED_DEFAULT_REPR
            if not _WARNED_DEFAULT_REPR:
                import warnings
                _WARNED_DEFAULT_REPR = True
                warnings.warn("keygen without transcript_repr: the key is hashed from this package's own description of "
                              "it, so the proofs verify under the in-repository verifier only; pass the scalar the Rust "
                              "side derives from the pinned key (plonk.rs:91-109) for interoperable proofs", stacklevel=3)
        self.transcript_repr = transcript_repr % R


_WARNED_DEFAULT_REPR = False


class ProvingKey:
    """plonk::ProvingKey: vk, l0 / l_last / l_active_row (extended), fixed_values / fixed_polys, the permutation
    proving key (sigma values and polynomials) and the Evaluator"""


def identity_mapping(m: int, n: int) -> np.ndarray:
    """permutation::keygen::Assembly::new: mapping[i][j] = (i, j), as an (m, n, 2) array"""
    out = np.empty((m, n, 2), dtype=np.int64)
    out[..., 0] = np.arange(m)[:, None]
    out[..., 1] = np.arange(n)[None, :]
    return out


def keygen(params, cs, fixed: np.ndarray, mapping, engine=None, zeta: Optional[int] = None,
           transcript_repr: Optional[int] = None) -> ProvingKey:
    """keygen_vk + keygen_pk for a laid-out circuit.  fixed: (num_fixed, n, 4) Lagrange columns; mapping: the
    permutation as (m, n, 2) integers (column position in cs.permutation_columns, row), Assembly.mapping."""
    from .domain import EvaluationDomain
    domain = EvaluationDomain(cs.degree(), params.k) if zeta is None else EvaluationDomain(cs.degree(), params.k, zeta)
    E = engine or Engine(params, domain)
    n = domain.n
    fixed = np.ascontiguousarray(fixed, dtype=np.uint64).reshape(cs.num_fixed, n, 4)
    m = len(cs.permutation_columns)
    mapping = np.asarray(mapping, dtype=np.int64).reshape(m, n, 2)
    # sigma_i[j] = delta^i' * omega^j' with (i', j') = mapping[i][j] (permutation/keygen.rs:196-233); the columns
    # delta^i * omega^j are the size-n transform of the polynomial delta^i * X
    ident = np.zeros((max(1, m), n, 4), dtype=np.uint64)
    d = 1
    for i in range(m):
        ident[i, 1] = _fr.to_mont(d)
        E.fft(ident[i])
        d = d * DELTA % R
    sigmas = np.ascontiguousarray(ident[mapping[..., 0], mapping[..., 1]]) if m else np.zeros((0, n, 4), np.uint64)
    pk = ProvingKey()
    fixed_commitments = E.commit_lagrange(fixed) if cs.num_fixed else []         # keygen.rs:288-291
    permutation_commitments = E.commit_lagrange(sigmas) if m else []             # permutation/keygen.rs:244-251
    pk.vk = VerifyingKey(cs, domain, fixed_commitments, permutation_commitments, transcript_repr)
    pk.fixed_values = fixed
    pk.fixed_polys = E.lagrange_to_coeff(fixed.copy()) if cs.num_fixed else fixed
    pk.sigmas = sigmas
    pk.sigma_polys = E.lagrange_to_coeff(sigmas.copy()) if m else sigmas
    # l0, l_blind, l_last over the extended domain (keygen.rs:398-431)
    bf = cs.blinding_factors()
    one = _fr.to_mont(1)
    lag = np.zeros((3, n, 4), dtype=np.uint64)
    lag[0, 0] = one
    lag[1, n - bf:] = one
    lag[2, n - bf - 1] = one
    ext = E.coeff_to_extended(E.lagrange_to_coeff(lag))
    pk.l0, l_blind, pk.l_last = ext[0], ext[1], ext[2]
    ones = np.broadcast_to(one, pk.l0.shape)
    pk.l_active_row = E.fr_vec("sub", ones, E.fr_vec("add", pk.l_last, l_blind))
    pk.ev = build_evaluator(cs)
    return pk


# --------------------------------------------------------------------------
# logup multiplicities (host: sort + binary search, as in the reference)
# --------------------------------------------------------------------------
def _sort_keys(canonical: np.ndarray) -> np.ndarray:
    """canonical (n, 4) little-endian limbs -> fixed-width big-endian byte strings, whose order is the integer order"""
    be = np.ascontiguousarray(canonical[:, ::-1]).astype(">u8")
    return np.frombuffer(be.tobytes(), dtype="S32")


def logup_multiplicity(inputs_canonical: Sequence[np.ndarray], table_canonical: np.ndarray, usable: int, n: int
                       ) -> np.ndarray:
    """logup/prover.rs:115-184 -> m as integer counts per row.  The table's usable rows are stably sorted by value
    and every input value is looked up with <[T]>::binary_search_by_key; when the table repeats a value the row that
    search returns gets the count, so the probe sequence of the pinned toolchain's implementation
    (mid = left + size / 2, first Equal probe wins) is reproduced on the run [lo, hi) of equal keys."""
    keys = _sort_keys(np.ascontiguousarray(table_canonical[:usable]))
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    m = np.zeros(n, dtype=np.int64)
    for inp in inputs_canonical:
        vals = _sort_keys(np.ascontiguousarray(inp[:usable]))
        lo = np.searchsorted(skeys, vals, side="left")
        hi = np.searchsorted(skeys, vals, side="right")
        if np.any(hi <= lo):
            raise B2Error(B2_ERR_ARG, "logup binary_search_by_key should hit")
        ulo, inverse = np.unique(lo, return_inverse=True)
        uhi = np.zeros_like(ulo)
        uhi[inverse] = hi
        found = np.full(ulo.shape, -1, dtype=np.int64)
        left = np.zeros_like(ulo)
        right = np.full_like(ulo, usable)
        size = right - left
        while np.any(found < 0):
            live = found < 0
            mid = left + size // 2
            less = live & (mid < ulo)
            greater = live & (mid >= uhi)
            hit = live & ~less & ~greater
            found[hit] = mid[hit]
            left = np.where(less, mid + 1, left)
            right = np.where(greater, mid, right)
            size = right - left
        m += np.bincount(order[found[inverse]], minlength=n)
    return m


class _DevArray:
    """raw device memory as a CUDA-array-interface object (so torch.distributed can move it without a copy)"""

    def __init__(self, ptr: int, shape, typestr: str = "<i8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def logup_multiplicity_device(inputs_ptr: int, n_inputs: int, table_ptr: int, usable: int, n: int, m_ptr: int) -> int:
    """logup_multiplicity on resident data (b2_logup_multiplicity_dev: radix sort of the table + the reference's binary
    search per input value, csrc/lookup.cuh).  inputs_ptr: n_inputs columns of n Montgomery-form field elements (the
    compressed inputs); table_ptr: the compressed table column; m_ptr receives n Montgomery-form elements: the counts in
    rows < usable, zeros above.  Returns the largest count."""
    import ctypes
    from ._lib import check, lib
    largest = ctypes.c_uint64()
    vp = ctypes.c_void_p
    check(lib().b2_logup_multiplicity_dev(vp(inputs_ptr), n_inputs, vp(table_ptr), usable, n, vp(m_ptr),
                                          ctypes.byref(largest)))
    return int(largest.value)


# --------------------------------------------------------------------------
# create_proof
# --------------------------------------------------------------------------
def create_proof(params, pk: ProvingKey, advice: np.ndarray, instances: Sequence[Sequence[int]], rng,
                 sign_bit: int = 7, engine=None, advice_max_bits: Optional[int] = None,
                 timings: Optional[dict] = None, use_gwc: bool = True) -> bytes:
    """plonk::create_proof (use_gwc=True, GWC multiopen, plonk/prover.rs:1759-1781) or create_proof_with_shplonk
    (use_gwc=False, :1737-1757) for one circuit instance, advice given (create_proof_from_witness).

    advice: (num_advice, n, 4) Lagrange columns in host memory (pinned memory makes the one upload faster); the
    blinding rows are written into it.  instances: per instance column the public values (canonical ints), zero
    padded internally.  advice_max_bits: the bound handed to commit_lagrange_with_bound; None = the engine scans each
    column for its largest scalar as the reference does (plonk/prover.rs:945-962, 296); a caller that knows its
    witness range passes it and gets the upload pipelined with the commitments.
    engine: ResidentEngine(params, domain) by default; pass one to keep the proving key resident across proofs.

    Random values come from `rng` in this order (vector draws):
      1. u16_vec(num_advice * (bf + 1)): advice column i takes [i*(bf+1), (i+1)*(bf+1)) for its last bf+1 rows
      2. per lookup: u16_vec(bf + 1), the last rows of m
      3. per permutation set: fr_vec(bf);  4. per lookup, per z: fr_vec(bf);  5. per shuffle group: fr_vec(bf)
      6. vanishing random polynomial: fr_vec(k) for `random`, u64_vec(4) = the 256-bit ChaCha20 key of the
         per-coefficient streams a, u, b, v (vanishing_streams): coeff[i] = (a_i + random[u_i % k]) * (b_i + random[v_i % k]),
         vanishing/prover.rs:48-63 -- the reference draws those four per coefficient from thread_rng (ChaCha too)
    """
    return create_proof_multi(params, pk, [advice], [instances], rng, sign_bit=sign_bit, engine=engine,
                              advice_max_bits=advice_max_bits, timings=timings, use_gwc=use_gwc)


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


def create_proof_multi(params, pk: ProvingKey, advices: Sequence[np.ndarray],
                       instances: Sequence[Sequence[Sequence[int]]], rng, sign_bit: int = 7, engine=None,
                       advice_max_bits: Optional[int] = None, timings: Optional[dict] = None,
                       use_gwc: bool = True) -> bytes:
    """ONE proof for `len(advices)` instances of the circuit: `create_proof_ext(circuits: &[ConcreteCircuit],
    instances: &[&[&[C::Scalar]]])`, plonk/prover.rs:206-222, with the advice of every instance given
    (create_proof_from_witness, :916-1500).  Every phase runs over all instances before the next challenge is drawn
    (instance commitments, advice commitments | theta | m commitments | beta, gamma | permutation z of every
    instance, lookup z of every instance, shuffle z of every instance | ...), evaluate_h folds the instances into one
    accumulator (evaluation.rs:839-845) and the multiopen queries are listed instance by instance (:1466-1512).
    Random draws: the order of create_proof, each numbered item once per instance before the next item."""
    import time
    vk = pk.vk
    cs, domain = vk.cs, vk.domain
    if len(advices) == 0 or len(advices) != len(instances):
        raise B2Error(B2_ERR_ARG, "InvalidInstances: one list of instance columns per circuit instance")
    from .evaluation import set_active_pool
    E = engine or ResidentEngine(params, domain)
    prev_pool = set_active_pool(getattr(E, "_pool", None))
    try:
        return _create_proof(E, pk, cs, domain, list(advices), list(instances), rng, sign_bit, advice_max_bits, timings,
                             time, use_gwc)
    finally:
        if engine is None:
            E.free()
        else:
            E.release()
        set_active_pool(prev_pool)


def _create_proof(E, pk, cs, domain, advices, instances, rng, sign_bit, advice_max_bits, timings, time, use_gwc) -> bytes:
    vk = pk.vk
    n, k = domain.n, domain.k
    bf = cs.blinding_factors()
    usable = n - (bf + 1)
    queries = cs.queries()
    tr = Blake2bWrite(sign_bit)
    t_last = [time.perf_counter()]
    C = len(advices)

    def lap(name: str) -> None:
        if timings is not None:
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_last[0]
            t_last[0] = now

    key = E.key_blocks(pk)

    # ---- create_single_instances (plonk/prover.rs:85-173)
    tr.common_scalar(vk.transcript_repr)
    inst_values, inst_polys = [], []
    for circuit_instances in instances:
        if len(circuit_instances) != cs.num_instance:
            raise B2Error(B2_ERR_ARG, "InvalidInstances")
        inst_host = np.zeros((cs.num_instance, n, 4), dtype=np.uint64)
        for i, values in enumerate(circuit_instances):
            if len(values) > usable:
                raise B2Error(B2_ERR_ARG, "InstanceTooLarge")
            if len(values):
                inst_host[i, :len(values)] = _mont_vec(values)
        instance_values = E.put(inst_host)
        if cs.num_instance:
            for c in E.commit_lagrange(instance_values, _fr.NUM_BITS):
                tr.common_point(c)
        instance_polys = E.copy(instance_values)
        if cs.num_instance:
            E.lagrange_to_coeff(instance_polys)
        inst_values.append(instance_values)
        inst_polys.append(instance_polys)
    lap("instance")

    # ---- advice (:964-1010)
    advs = []
    for advice in advices:
        if advice.shape != (cs.num_advice, n, 4) or advice.dtype != np.uint64 or not advice.flags.c_contiguous:
            raise B2Error(B2_ERR_ARG, f"advice must be a C-contiguous uint64 array of shape ({cs.num_advice}, {n}, 4)")
        blind = rng.u16_vec(cs.num_advice * (bf + 1))
        advice[:, usable:] = _mont_vec(blind).reshape(cs.num_advice, bf + 1, 4)
        adv, points = E.put_and_commit_lagrange(advice, advice_max_bits)
        for c in points:
            tr.write_point(c)
        advs.append(adv)
    theta = tr.squeeze_challenge()
    lap("advice")

    # ---- lookups: compress, multiplicities, m commitments (:334-366, logup/prover.rs:70-256)
    n_lookups = len(cs.lookups)
    all_ms = []
    for adv, instance_values in zip(advs, inst_values):
        ms, m_bits = E.multiplicity_block(cs, pk, adv, instance_values, theta,
                                          [rng.u16_vec(bf + 1) for _ in range(n_lookups)])
        all_ms.append((ms, m_bits))
    for ms, m_bits in all_ms:
        if n_lookups:
            for c in E.commit_lagrange(ms, m_bits):
                tr.write_point(c)
    all_ms = [ms for ms, _ in all_ms]
    beta = tr.squeeze_challenge()
    gamma = tr.squeeze_challenge()
    lap("lookup_m")

    # ---- z columns (:411-633): permutation, lookups, shuffles -- built where the engine keeps its columns,
    # blinded, then committed and brought to coefficient form as ONE batch per instance; the points go to the
    # transcript in the reference's order: permutation z of every instance, then lookup z, then shuffle z
    chunk_len = cs.degree() - 2
    n_perm = (len(cs.permutation_columns) + chunk_len - 1) // chunk_len
    lookup_z_counts = [len(lk["input_expressions_sets"]) for lk in cs.lookups]
    n_shuffles = len(cs.shuffles)
    n_z = n_perm + sum(lookup_z_counts) + n_shuffles
    z_blocks = [E.alloc(n_z) for _ in range(C)]
    for ci in range(C):
        if n_perm:
            blinds = [rng.fr_vec(bf) for _ in range(n_perm)]
            E.permutation_z(cs, pk, advs[ci], inst_values[ci], beta, gamma, blinds, E.cols(z_blocks[ci])[:n_perm])
    # A multi-GPU engine builds only the z columns it will commit (owns_columns) and receives the others, already in
    # coefficient form, inside commit_lagrange_and_ifft; the z columns of one lookup are chained (z_i starts where
    # z_(i-1) ended), so a lookup is built whole by every rank that owns one of them.  The random draws do not depend
    # on ownership: every rank consumes the same stream.
    owns = getattr(E, "owns_columns", None) or (lambda block, lo, hi: True)
    for ci in range(C):
        zc, m_cols, pos = E.cols(z_blocks[ci]), E.cols(all_ms[ci]), n_perm
        for li, lk in enumerate(cs.lookups):
            cnt = lookup_z_counts[li]
            mine = owns(z_blocks[ci], pos, pos + cnt)
            if mine:
                E.logup_z(cs, lk, pk, advs[ci], inst_values[ci], m_cols[li], theta, beta, zc[pos:pos + cnt])
            for z in zc[pos:pos + cnt]:
                blind = rng.fr_vec(bf)
                if mine:
                    E.write_rows(z, n - bf, blind)
            pos += cnt
    for ci in range(C):
        zc, pos = E.cols(z_blocks[ci]), n_perm + sum(lookup_z_counts)
        for gi, group in enumerate(cs.shuffles):
            blind = rng.fr_vec(bf)
            if owns(z_blocks[ci], pos + gi, pos + gi + 1):
                E.shuffle_z(cs, group, pk, advs[ci], inst_values[ci], theta, beta, zc[pos + gi])
                E.write_rows(zc[pos + gi], n - bf, blind)
    z_points = [E.commit_lagrange_and_ifft(zb) if n_z else [] for zb in z_blocks]
    for pts in z_points:
        for c in pts[:n_perm]:
            tr.write_point(c)
    for pts in z_points:
        for c in pts[n_perm:n_perm + sum(lookup_z_counts)]:
            tr.write_point(c)
    for pts in z_points:
        for c in pts[n_perm + sum(lookup_z_counts):]:
            tr.write_point(c)
    per_circuit = []
    for ci in range(C):
        if n_lookups:
            E.lagrange_to_coeff(all_ms[ci])                                  # lagrange_to_coeff_st(l.0), :497
        zc, m_cols = E.cols(z_blocks[ci]), E.cols(all_ms[ci])
        lookups, pos = [], n_perm
        for li, cnt in enumerate(lookup_z_counts):
            lookups.append({"z": zc[pos:pos + cnt], "m": m_cols[li]})
            pos += cnt
        per_circuit.append({"perm": zc[:n_perm], "lookups": lookups, "shuffles": zc[pos:pos + n_shuffles]})
    lap("z_columns")

    # ---- vanishing commit, y (:635-639, vanishing/prover.rs:41-70)
    random = rng.fr_vec(k)
    random_block = E.random_poly(random, np.asarray(rng.u64_vec(4), dtype="<u8").tobytes())
    random_poly = E.cols(random_block)[0]
    tr.write_point(E.commit(random_block)[0])
    y = tr.squeeze_challenge()
    lap("vanishing_commit")

    # ---- h(X) (:640-690, vanishing/prover.rs:64-110).  The reference folds every instance into one accumulator:
    # after instance j it holds V_0 y^(jT) + ... + V_j, V_j = the fold of instance j alone from zero, T = fold_steps.
    # Division by the vanishing polynomial and extended_to_coeff are linear, so the pieces of h are the same Horner
    # fold (in y^T) of the pieces each instance yields on its own.
    h_blocks = []
    for ci in range(C):
        E.lagrange_to_coeff(advs[ci])                                        # lagrange_to_coeff_st per column, :643-646
        h_blocks.append(E.evaluate_h_blocks(pk, advs[ci], inst_polys[ci], z_blocks[ci], all_ms[ci], n_perm,
                                            lookup_z_counts, n_shuffles, y, beta, gamma, theta))
    if C == 1:
        h_block = h_blocks[0]
    else:
        y_t = pow(y, fold_steps(cs), R)
        pieces = [E.cols(hb) for hb in h_blocks]
        h_block = E.stack([E.poly_combine([pieces[ci][p] for ci in range(C)], y_t) for p in range(len(pieces[0]))])
    for c in E.commit(h_block):
        tr.write_point(c)
    x = tr.squeeze_challenge()
    xn = pow(x, n, R)
    lap("h_poly")

    # ---- evaluations (:694-790)
    rot = lambda at: x * pow(domain._omega if at >= 0 else domain._omega_inv, abs(at), R) % R      # noqa: E731
    known: Dict[Tuple[int, int], int] = {}

    def ev(poly, point: int) -> int:
        """eval_polynomial, remembered: SHPLONK asks for the same (polynomial, point) pairs again"""
        key = (_poly_key(poly), point)
        if key not in known:
            known[key] = E.eval_polynomial(poly, point)
        return known[key]
    all_advice_polys = [E.cols(adv) for adv in advs]
    all_inst_cols = [E.cols(ip) for ip in inst_polys]
    fixed_polys, sigma_polys = E.cols(key["fixed_polys"]), E.cols(key["sigma_polys"])
    # The evaluations are written in the reference's order, but no challenge is drawn between them: list the
    # (polynomial, point) pairs first, evaluate them with ONE engine call per distinct point, then write.
    wanted: List[Tuple[object, int]] = []
    for inst_cols in all_inst_cols:
        for col, at in queries["Instance"]:
            wanted.append((inst_cols[col], rot(at)))
    for advice_polys in all_advice_polys:
        for col, at in queries["Advice"]:
            wanted.append((advice_polys[col], rot(at)))
    for col, at in queries["Fixed"]:
        wanted.append((fixed_polys[col], rot(at)))
    h_poly = E.poly_combine(list(reversed(E.cols(h_block))), xn)       # fold acc * xn + piece over rev pieces
    wanted.append((random_poly, x))
    for poly in sigma_polys:
        wanted.append((poly, x))
    last = -(bf + 1)
    x_next, x_last = rot(1), rot(last)

    def eval_z_set(polys):
        for i, z in enumerate(polys):
            wanted.append((z, x))
            wanted.append((z, x_next))
            if i + 1 < len(polys):
                wanted.append((z, x_last))

    for pc in per_circuit:
        eval_z_set(pc["perm"])
    for pc in per_circuit:
        for lk in pc["lookups"]:
            wanted.append((lk["m"], x))
            eval_z_set(lk["z"])
    for pc in per_circuit:
        for z in pc["shuffles"]:
            wanted.append((z, x))
            wanted.append((z, x_next))
    eval_many = getattr(E, "eval_polynomials", None)
    if eval_many is not None:
        by_point: Dict[int, list] = {}
        for poly, point in wanted:
            if (_poly_key(poly), point) not in known:
                known[(_poly_key(poly), point)] = None
                by_point.setdefault(point, []).append(poly)
        for point, polys in by_point.items():
            for poly, value in zip(polys, eval_many(polys, point)):
                known[(_poly_key(poly), point)] = value
    for poly, point in wanted:
        tr.write_scalar(ev(poly, point))
    lap("evaluations")

    # ---- multiopen queries (:792-838) as (rotation, point, polynomial), instance by instance
    qs: List[Tuple[int, int, object]] = []

    def open_z_set(polys):
        for z in polys:
            qs.append((0, x, z))
            qs.append((1, x_next, z))
        for z in list(reversed(polys))[1:]:
            qs.append((last, x_last, z))

    for ci, pc in enumerate(per_circuit):
        for col, at in queries["Instance"]:
            qs.append((at, rot(at), all_inst_cols[ci][col]))
        for col, at in queries["Advice"]:
            qs.append((at, rot(at), all_advice_polys[ci][col]))
        open_z_set(pc["perm"])
        for lk in pc["lookups"]:
            qs.append((0, x, lk["m"]))
            open_z_set(lk["z"])
        for z in pc["shuffles"]:
            qs.append((0, x, z))
            qs.append((1, x_next, z))
    for col, at in queries["Fixed"]:
        qs.append((at, rot(at), fixed_polys[col]))
    for poly in sigma_polys:
        qs.append((0, x, poly))
    qs.append((0, x, h_poly))
    qs.append((0, x, random_poly))

    if use_gwc:
        gwc_create_proof(E, tr, qs)
    else:
        shplonk_create_proof(E, tr, qs, evaluate=ev)
    lap("multiopen")
    return tr.finalize()


def create_proof_with_shplonk(params, pk, advice, instances, rng, **kw) -> bytes:
    """plonk/prover.rs:1737-1757"""
    return create_proof(params, pk, advice, instances, rng, use_gwc=False, **kw)


def gwc_create_proof(E, tr: Blake2bWrite, queries) -> None:
    """poly/multiopen/gwc/prover.rs:19-173.  construct_intermediate_sets (gwc.rs:38-62) groups the queries by
    ROTATION (BTreeMap<Rotation, _>, ascending), point = the first query's; per group the polynomials are folded with
    v by the engine (the reference's cuda build does the same for groups of more than four), the folded polynomial
    is evaluated and divided by (X - z) there, and all witness polynomials are committed as one batch."""
    v = tr.squeeze_challenge()
    groups: Dict[int, list] = {}
    for q in queries:
        groups.setdefault(q[0], []).append(q)
    witnesses = []
    for r in sorted(groups):
        group = groups[r]
        z = group[0][1]
        if any(q[1] != z for q in group):
            raise B2Error(B2_ERR_ARG, "assert_eq!(query.get_point(), z)")
        poly_batch = E.poly_combine([q[2] for q in group], v)
        eval_batch = E.eval_polynomial(poly_batch, z)
        E.sub_constant(poly_batch, eval_batch)
        witnesses.append(E.kate_division_padded(poly_batch, z))
    for c in E.commit(E.stack(witnesses)):
        tr.write_point(c)


# --------------------------------------------------------------------------
# SHPLONK (poly/multiopen/shplonk.rs, shplonk/prover.rs)
# --------------------------------------------------------------------------
def _poly_key(col) -> int:
    """identity of a polynomial handle (PolynomialPointer compares addresses, poly/multiopen.rs:150-158)"""
    return col.ptr if isinstance(col, DevBlock) else col.ctypes.data


def lagrange_interpolate(points: Sequence[int], evals: Sequence[int]) -> List[int]:
    """arithmetic.rs:848-906 on a handful of points (host scalars): coefficients of the interpolant"""
    m = len(points)
    coeffs = [0] * m
    for j in range(m):
        num, den = [1], 1
        for i in range(m):
            if i == j:
                continue
            nxt = [0] * (len(num) + 1)
            for t, c in enumerate(num):
                nxt[t] = (nxt[t] - c * points[i]) % R
                nxt[t + 1] = (nxt[t + 1] + c) % R
            num = nxt
            den = den * (points[j] - points[i]) % R
        scale = evals[j] * _fr.inv(den) % R
        for t, c in enumerate(num):
            coeffs[t] = (coeffs[t] + c * scale) % R
    return coeffs


def _vanishing_at(roots: Sequence[int], z: int) -> int:
    """evaluate_vanishing_polynomial, arithmetic.rs:905-923"""
    acc = 1
    for p in roots:
        acc = (z - p) * acc % R
    return acc


def _eval_small(coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def shplonk_create_proof(E, tr: Blake2bWrite, queries, evaluate=None) -> None:
    """poly/multiopen/shplonk/prover.rs:78-234 over the engine.  The per-commitment subtractions of the reference
    (P - R for the quotient, P - R(u) for the linearisation) are linear, so each rotation set folds its polynomials
    with y ONCE on the device (F = sum y^j P_j) and the low-degree parts are handled as a few host scalars:
    N = F - R_fold (first |points| coefficients adjusted in place), Q = N / Z by repeated kate_division,
    L = N + R_fold - r_fold(u).  Scaling by z_i is the Horner fold of [L, 0] at z_i."""
    evaluate = evaluate or E.eval_polynomial       # the prover passes its memo of the evaluations it already wrote
    y = tr.squeeze_challenge()
    # construct_intermediate_sets, shplonk.rs:57-150
    point_of: Dict[int, int] = {}
    for q in queries:
        if point_of.setdefault(q[0], q[1]) != q[1]:
            raise B2Error(B2_ERR_ARG, "assert_eq!(*point, query.get_point())")
    super_point_set = [point_of[r] for r in sorted(point_of)]
    order, rots = [], {}
    for q in queries:
        k = _poly_key(q[2])
        if k not in rots:
            rots[k] = set()
            order.append((k, q[2]))
        rots[k].add(q[0])
    groups: Dict[tuple, list] = {}
    for k, poly in order:
        groups.setdefault(tuple(sorted(rots[k])), []).append(poly)
    sets = []
    for rs in sorted(groups):                                  # BTreeMap<BTreeSet<Rotation>, _>: lexicographic
        points = [point_of[r] for r in rs]
        polys = groups[rs]
        lows = [lagrange_interpolate(points, [evaluate(p, pt) for pt in points]) for p in polys]            # :33-48
        sets.append((polys, points, lows))
    v = tr.squeeze_challenge()

    def fold_small(vals: Sequence[Sequence[int]], c: int) -> List[int]:
        acc = [0] * len(vals[0])
        for p in vals:
            acc = [(a * c + b) % R for a, b in zip(acc, p)]
        return acc

    numerators, quotients = [], []
    for polys, points, lows in sets:                           # quotient_contribution, :97-128
        f = E.poly_combine(polys, y)
        r_fold = fold_small(lows, y)
        E.sub_low_degree(f, r_fold)                            # N_i = sum y^j (P_j - R_j)
        q = f
        for p in points:                                       # div_by_vanishing, :20-26
            q = E.kate_division_padded(q, p)
        numerators.append((f, r_fold))
        quotients.append(q)
    h_x = E.poly_combine(quotients, v)
    tr.write_point(E.commit(E.stack([h_x]))[0])
    u = tr.squeeze_challenge()
    zt_eval = _vanishing_at(super_point_set, u)
    lins, z_diffs = [], []
    for (polys, points, lows), (f, r_fold) in zip(sets, numerators):     # linearisation_contribution, :163-192
        z_i = _vanishing_at([p for p in super_point_set if p not in points], u)
        c_i = fold_small([[_eval_small(low, u)] for low in lows], y)[0]
        adjust = [(-a) % R for a in r_fold]
        adjust[0] = (adjust[0] + c_i) % R
        E.sub_low_degree(f, adjust)                            # L_i = N_i + R_fold - r_fold(u)
        lins.append(E.scale(f, z_i))
        z_diffs.append(z_i)
    l_x = E.sub_cols(E.poly_combine(lins, v), E.scale(h_x, zt_eval))     # :205-211
    if E.eval_polynomial(l_x, u) != 0:                         # sanity check, :213-216
        raise B2Error(B2_ERR_ARG, "shplonk: linearisation polynomial does not vanish at u")
    h2 = E.scale(E.kate_division_padded(l_x, u), _fr.inv(z_diffs[0]))    # :218-224
    tr.write_point(E.commit(E.stack([h2]))[0])
