"""CPU oracle for the quotient-evaluation path.

TEST INFRASTRUCTURE ONLY (same rules as oracle/bn254.py): plain Python big-int restatement of
  * Evaluator::new / add_expression            plonk/evaluation.rs:307-448, 623-776
  * Calculation::evaluate / ValueSource::get   plonk/evaluation.rs:60-268
  * Evaluator::evaluate_h (non-cuda variant)   plonk/evaluation.rs:778-1226
  * evaluate_with_theta                        plonk/evaluation.rs:2330-2398
  * permutation::Argument::commit              plonk/permutation/prover.rs:47-165
  * permutation keygen (sigma polynomials)     plonk/permutation/keygen.rs:196-262
  * logup compress / commit_z                  plonk/logup/prover.rs:63-420
  * shuffle commit_product                     plonk/shuffle/prover.rs:107-157
  * l0 / l_last / l_active_row                 plonk/keygen.rs:398-431
All paths relative to /root/reference/halo2_proofs/src.  Values are canonical integers mod r
(not Montgomery); oracle/bn254.py converts at the ABI boundary.

PARITY STATUS: byte-level parity unpinned (see oracle/bn254.py).  What pins this file
(tests/test_oracle_plonk.py): for a satisfied circuit the evaluate_h output must be divisible by
the vanishing polynomial with a quotient of degree < n * (degree - 1) -- the identity the
reference's prove -> verify tests rest on -- and it must stop being divisible when one witness
cell, one copy constraint, one lookup input or one shuffle cell is broken.
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence, Tuple

from .bn254 import FR_GENERATOR, FR_S, R_MOD, EvaluationDomain, fr_inv

FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)  # Fr::DELTA = g^(2^S)

# --------------------------------------------------------------------------
# Expression trees (plonk/circuit.rs Expression): plain tuples
# --------------------------------------------------------------------------
def Const(v: int): return ("Constant", v % R_MOD)
def Fixed(c: int, rot: int = 0): return ("Fixed", c, rot)
def Advice(c: int, rot: int = 0): return ("Advice", c, rot)
def Instance(c: int, rot: int = 0): return ("Instance", c, rot)
def Neg(e): return ("Negated", e)
def Sum(a, b): return ("Sum", a, b)
def Sub(a, b): return ("Sum", a, ("Negated", b))   # Expression: a - b is stored as a + (-b)
def Prod(a, b): return ("Product", a, b)
def Scaled(e, f: int): return ("Scaled", e, f % R_MOD)


def expr_degree(e) -> int:
    t = e[0]
    if t == "Constant":
        return 0
    if t in ("Fixed", "Advice", "Instance"):
        return 1
    if t in ("Negated", "Scaled"):
        return expr_degree(e[1])
    if t == "Sum":
        return max(expr_degree(e[1]), expr_degree(e[2]))
    return expr_degree(e[1]) + expr_degree(e[2])


def eval_expr(e, row: int, n: int, rot_scale: int, fixed, advice, instance) -> int:
    """Expression::evaluate with the closures of evaluate_with_theta (evaluation.rs:2350-2385)."""
    t = e[0]
    if t == "Constant":
        return e[1]
    if t in ("Fixed", "Advice", "Instance"):
        cols = {"Fixed": fixed, "Advice": advice, "Instance": instance}[t]
        return cols[e[1]][(row + e[2] * rot_scale) % n]
    if t == "Negated":
        return (-eval_expr(e[1], row, n, rot_scale, fixed, advice, instance)) % R_MOD
    if t == "Sum":
        return (eval_expr(e[1], row, n, rot_scale, fixed, advice, instance)
                + eval_expr(e[2], row, n, rot_scale, fixed, advice, instance)) % R_MOD
    if t == "Product":
        return (eval_expr(e[1], row, n, rot_scale, fixed, advice, instance)
                * eval_expr(e[2], row, n, rot_scale, fixed, advice, instance)) % R_MOD
    if t == "Scaled":
        return eval_expr(e[1], row, n, rot_scale, fixed, advice, instance) * e[2] % R_MOD
    raise ValueError(t)


def evaluate_with_theta(expressions, n: int, rot_scale: int, fixed, advice, instance, theta: int) -> List[int]:
    """evaluation.rs:2330-2398: fold of the expressions with theta, per row."""
    out = [0] * n
    for row in range(n):
        acc = 0
        for e in expressions:
            acc = (acc * theta + eval_expr(e, row, n, rot_scale, fixed, advice, instance)) % R_MOD
        out[row] = acc
    return out


# --------------------------------------------------------------------------
# Evaluator (GraphEvaluator-style CSE): evaluation.rs:270-776
# --------------------------------------------------------------------------
_VS_RANK = {"Constant": 0, "Intermediate": 1, "Fixed": 2, "Advice": 3, "Instance": 4}


def _vs_key(v):
    """derive(PartialOrd) on ValueSource: variant order, then fields (evaluation.rs:45-58)."""
    return (_VS_RANK[v[0]],) + tuple(v[1:])


class ConstraintSystem:
    """The part of plonk::ConstraintSystem that evaluate_h reads."""

    def __init__(self, num_fixed: int, num_advice: int, num_instance: int, degree: int, blinding_factors: int = 5):
        self.num_fixed, self.num_advice, self.num_instance = num_fixed, num_advice, num_instance
        self.gates: List[List[tuple]] = []           # gate.polynomials()
        self.lookups: List[dict] = []                # {"table_expressions": [...], "input_expressions_sets": [[[...], ...], ...]}
        self.shuffles: List[List[dict]] = []         # groups of {"input_expressions": [...], "shuffle_expressions": [...]}
        self.permutation_columns: List[Tuple[str, int]] = []   # ("Advice"|"Fixed"|"Instance", index)
        self._degree = degree
        self._blinding_factors = blinding_factors

    def degree(self) -> int:
        return self._degree

    def blinding_factors(self) -> int:
        return self._blinding_factors


class Evaluator:
    def __init__(self):
        self.constants: List[int] = []
        self.rotations: List[int] = []
        self.calculations: List[tuple] = []
        self.value_parts: List[tuple] = []
        self.lookup_results: List[Tuple[tuple, List[tuple], List[tuple]]] = []
        self.shuffle_results: List[Tuple[tuple, tuple]] = []

    # :623-633
    def add_rotation(self, rotation: int) -> int:
        if rotation in self.rotations:
            return self.rotations.index(rotation)
        self.rotations.append(rotation)
        return len(self.rotations) - 1

    # :635-647
    def add_constant(self, c: int):
        c %= R_MOD
        if c in self.constants:
            return ("Constant", self.constants.index(c))
        self.constants.append(c)
        return ("Constant", len(self.constants) - 1)

    # :650-668
    def add_calculation(self, calc: tuple):
        if calc in self.calculations:
            return ("Intermediate", self.calculations.index(calc))
        self.calculations.append(calc)
        return ("Intermediate", len(self.calculations) - 1)

    # :671-776
    def add_expression(self, e):
        t = e[0]
        if t == "Constant":
            return self.add_constant(e[1])
        if t in ("Fixed", "Advice", "Instance"):
            rot_idx = self.add_rotation(e[2])
            return self.add_calculation(("Store", (t, e[1], rot_idx)))
        if t == "Negated":
            if e[1][0] == "Constant":
                return self.add_constant(-e[1][1])
            ra = self.add_expression(e[1])
            if ra == ("Constant", 0):
                return ra
            return self.add_calculation(("Negate", ra))
        if t == "Sum":
            if e[2][0] == "Negated":
                ra = self.add_expression(e[1])
                rb = self.add_expression(e[2][1])
                if ra == ("Constant", 0):
                    return rb     # (sic) :727-728 returns result_b, not its negation
                if rb == ("Constant", 0):
                    return ra
                return self.add_calculation(("Sub", ra, rb))
            ra = self.add_expression(e[1])
            rb = self.add_expression(e[2])
            if ra == ("Constant", 0):
                return rb
            if rb == ("Constant", 0):
                return ra
            if _vs_key(ra) <= _vs_key(rb):
                return self.add_calculation(("Add", ra, rb))
            return self.add_calculation(("Add", rb, ra))
        if t == "Product":
            ra = self.add_expression(e[1])
            rb = self.add_expression(e[2])
            if ra == ("Constant", 0) or rb == ("Constant", 0):
                return ("Constant", 0)
            if ra == ("Constant", 1):
                return rb
            if rb == ("Constant", 1):
                return ra
            if _vs_key(ra) <= _vs_key(rb):
                return self.add_calculation(("Mul", ra, rb))
            return self.add_calculation(("Mul", rb, ra))
        if t == "Scaled":
            if e[2] == 0:
                return ("Constant", 0)
            if e[2] == 1:
                return self.add_expression(e[1])
            cst = self.add_constant(e[2])
            ra = self.add_expression(e[1])
            return self.add_calculation(("Mul", ra, cst))
        raise ValueError(t)

    @classmethod
    def new(cls, cs: ConstraintSystem) -> "Evaluator":
        """:309-448 (the CPU structures; the gpu_* expression trees are not restated)."""
        ev = cls()
        ev.add_constant(0)
        constant_one = ev.add_constant(1)
        for gate in cs.gates:                                         # :317-323
            for poly in gate:
                ev.value_parts.append(ev.add_expression(poly))

        def evaluate_lc(expressions):                                  # :350-360
            parts = [ev.add_expression(x) for x in expressions]
            lc = parts[0]
            for part in parts[1:]:
                lc = ev.add_calculation(("LcTheta", lc, part))
            return lc

        def evaluate_compress_challenge(expressions):                  # :362-369
            return ev.add_calculation(("AddChallenge", evaluate_lc(expressions), "Beta"))

        for lookup in cs.lookups:                                      # :377-438
            table = ("AddChallenge", evaluate_lc(lookup["table_expressions"]), "Beta")
            input_sets = [[evaluate_compress_challenge(inp) for inp in s] for s in lookup["input_expressions_sets"]]
            products = []
            for cosets in input_sets:
                lc_product = cosets[0]
                for p in cosets[1:]:
                    lc_product = ev.add_calculation(("Mul", lc_product, p))
                products.append(("Store", lc_product))
            sums = []
            for cosets in input_sets:
                if len(cosets) > 1:
                    prods = []
                    for i in range(len(cosets)):
                        others = [v for j, v in enumerate(cosets) if j != i]
                        acc = others[0]
                        for v in others[1:]:
                            acc = ev.add_calculation(("Mul", acc, v))
                        prods.append(acc)
                    acc = prods[0]
                    for v in prods[1:]:
                        acc = ev.add_calculation(("Add", acc, v))
                    sums.append(("Store", acc))
                else:
                    sums.append(("Store", constant_one))
            ev.lookup_results.append((table, products, sums))

        for group in cs.shuffles:                                      # :539-578
            pairs = [(evaluate_lc(a["input_expressions"]), evaluate_lc(a["shuffle_expressions"])) for a in group]  # :536-548
            inputs = [p[0] for p in pairs]
            shuffles = [p[1] for p in pairs]
            product_inputs = ("AddChallenge", inputs[0], "Beta")
            for i, part in list(enumerate(inputs))[1:]:
                product_inputs = ("LcChallenge", part, ev.add_calculation(product_inputs), "Beta", i + 1)
            product_shuffles = ("AddChallenge", shuffles[0], "Beta")
            for i, part in list(enumerate(shuffles))[1:]:
                product_shuffles = ("LcChallenge", part, ev.add_calculation(product_shuffles), "Beta", i + 1)
            ev.shuffle_results.append((product_inputs, product_shuffles))
        return ev


def _get(src, rotations, constants, intermediates, fixed, advice, instance) -> int:
    """ValueSource::get, :60-85"""
    t = src[0]
    if t == "Constant":
        return constants[src[1]]
    if t == "Intermediate":
        return intermediates[src[1]]
    cols = {"Fixed": fixed, "Advice": advice, "Instance": instance}[t]
    return cols[src[1]][rotations[src[2]]]


def calc_evaluate(calc, rotations, constants, intermediates, fixed, advice, instance, beta, gamma, theta) -> int:
    """Calculation::evaluate, :115-267"""
    g = lambda s: _get(s, rotations, constants, intermediates, fixed, advice, instance)  # noqa: E731
    t = calc[0]
    if t == "Add":
        return (g(calc[1]) + g(calc[2])) % R_MOD
    if t == "Sub":
        return (g(calc[1]) - g(calc[2])) % R_MOD
    if t == "Mul":
        return g(calc[1]) * g(calc[2]) % R_MOD
    if t == "Negate":
        return (-g(calc[1])) % R_MOD
    if t == "LcChallenge":
        x = beta if calc[3] == "Beta" else gamma
        p = calc[4]
        xp = pow(x, p, R_MOD) if p > 1 else x
        return (g(calc[1]) + xp) * g(calc[2]) % R_MOD
    if t == "LcTheta":
        return (g(calc[1]) * theta + g(calc[2])) % R_MOD
    if t == "AddChallenge":
        x = beta if calc[2] == "Beta" else gamma
        return (g(calc[1]) + x) % R_MOD
    if t == "Store":
        return g(calc[1])
    raise ValueError(t)


# --------------------------------------------------------------------------
# keygen pieces
# --------------------------------------------------------------------------
def lagrange_basis_cosets(cs: ConstraintSystem, domain: EvaluationDomain):
    """keygen.rs:398-431 -> (l0, l_last, l_active_row) over the extended domain"""
    n, bf = domain.n, cs.blinding_factors()
    l0 = [0] * n
    l0[0] = 1
    l_blind = [0] * n
    for i in range(n - bf, n):
        l_blind[i] = 1
    l_last = [0] * n
    l_last[n - bf - 1] = 1
    ext = lambda v: domain.coeff_to_extended(domain.lagrange_to_coeff(v))  # noqa: E731
    l0e, l_blind_e, l_last_e = ext(l0), ext(l_blind), ext(l_last)
    l_active = [(1 - (a + b)) % R_MOD for a, b in zip(l_last_e, l_blind_e)]
    return l0e, l_last_e, l_active


def permutation_sigmas(cs: ConstraintSystem, domain: EvaluationDomain, mapping) -> List[List[int]]:
    """permutation/keygen.rs:196-233: sigma_i[j] = delta^(i') * omega^(j'), (i', j') = mapping[i][j]."""
    n = domain.n
    m = len(cs.permutation_columns)
    delta_omegas = []
    d = 1
    for _ in range(m):
        row, cur = [], d
        for _ in range(n):
            row.append(cur)
            cur = cur * domain.omega % R_MOD
        delta_omegas.append(row)
        d = d * FR_DELTA % R_MOD
    return [[delta_omegas[mapping[i][j][0]][mapping[i][j][1]] for j in range(n)] for i in range(m)]


def identity_mapping(m: int, n: int):
    return [[(i, j) for j in range(n)] for i in range(m)]


def mapping_copy(mapping, a: Tuple[int, int], b: Tuple[int, int]) -> None:
    """Join the cycles of cells a and b (permutation/keygen.rs Assembly::copy, restated as a cycle swap)."""
    # already in the same cycle?
    cur = mapping[a[0]][a[1]]
    while cur != a:
        if cur == b:
            return
        cur = mapping[cur[0]][cur[1]]
    if a == b:
        return
    mapping[a[0]][a[1]], mapping[b[0]][b[1]] = mapping[b[0]][b[1]], mapping[a[0]][a[1]]


def _column(kind_index, fixed, advice, instance):
    kind, idx = kind_index
    return {"Fixed": fixed, "Advice": advice, "Instance": instance}[kind][idx]


def batch_invert(v: List[int]) -> None:
    """ff::BatchInvert / arithmetic.rs batch_invert: zeros stay zero."""
    for i, x in enumerate(v):
        v[i] = fr_inv(x) if x else 0


# --------------------------------------------------------------------------
# grand products / sums (prover side)
# --------------------------------------------------------------------------
def permutation_commit(cs, domain, sigmas, advice, fixed, instance, beta, gamma, rng: random.Random) -> List[List[int]]:
    """permutation/prover.rs:47-165 -> the z polynomials (Lagrange basis), one per column chunk."""
    n = domain.n
    chunk_len = cs.degree() - 2
    bf = cs.blinding_factors()
    cols = cs.permutation_columns
    raw_zs = []
    for ci in range(0, len(cols), chunk_len):
        columns = cols[ci:ci + chunk_len]
        perms = sigmas[ci:ci + chunk_len]
        delta_omega = pow(FR_DELTA, ci, R_MOD)        # :77: DELTA^(i * chunk_len), i = chunk index
        modified = [1] * n
        for column, sigma in zip(columns, perms):    # :90-101
            values = _column(column, fixed, advice, instance)
            for i in range(n):
                modified[i] = modified[i] * ((beta * sigma[i] + gamma + values[i]) % R_MOD) % R_MOD
        batch_invert(modified)                        # :104
        for column in columns:                        # :108-121
            values = _column(column, fixed, advice, instance)
            for i in range(n):
                modified[i] = modified[i] * ((delta_omega * beta + gamma + values[i]) % R_MOD) % R_MOD
                delta_omega = delta_omega * domain.omega % R_MOD
            delta_omega = delta_omega * FR_DELTA % R_MOD
        z = [0] * n                                   # :135-141
        for i in range(1, n):
            z[i] = modified[i - 1]
        raw_zs.append(z)
    sets = []
    last_z = 1                                        # :146
    for z in raw_zs:
        z[0] = last_z                                 # :149-152
        for i in range(n - 1):
            z[i + 1] = z[i] * z[i + 1] % R_MOD
        for i in range(n - bf, n):                    # :156-158
            z[i] = rng.randrange(R_MOD)
        last_z = z[n - (bf + 1)]                      # :160
        sets.append(z)
    return sets


def logup_compress(cs, domain, lookup, theta, advice, fixed, instance, rng: random.Random):
    """logup/prover.rs:70-256 -> (compressed inputs per set, compressed table, multiplicities)."""
    n = domain.n
    bf = cs.blinding_factors()
    usable = n - bf - 1
    comp = lambda ex: evaluate_with_theta(ex, n, 1, fixed, advice, instance, theta)  # noqa: E731
    input_sets = [[comp(inp) for inp in s] for s in lookup["input_expressions_sets"]]
    table = comp(lookup["table_expressions"])
    # the reference sorts (value, index) and binary-searches: any row holding the value may win; the
    # multiplicity is credited to ONE row per distinct value (:115-172).  Parity of m therefore needs
    # the same tie rule: restated as "the row the stable sort + binary search lands on"; for tables
    # without repeated values (all that the tests use) the row is unique.
    first_row: Dict[int, int] = {}
    for i in range(usable):
        first_row.setdefault(table[i], i)
    m = [0] * n
    for s in input_sets:
        for inp in s:
            for v in inp[:usable]:
                m[first_row[v]] += 1                  # KeyError = "logup binary_search_by_key should hit"
    for i in range(usable, n):                        # :232-236: u16 blinding of m
        m[i] = rng.randrange(1 << 16)
    return input_sets, table, m


def logup_commit_z(cs, domain, input_sets, table, m, beta) -> List[List[int]]:
    """logup/prover.rs:263-336 -> raw z vectors (length n - blinding_factors), one per input set."""
    n = domain.n
    bf = cs.blinding_factors()
    base = [0] * n
    for inp in input_sets[0]:
        fi = [(beta + v) % R_MOD for v in inp]
        batch_invert(fi)
        base = [(a + b) % R_MOD for a, b in zip(base, fi)]
    ts = [(beta + v) % R_MOD for v in table]
    batch_invert(ts)
    base = [(s - t * mm) % R_MOD for s, t, mm in zip(base, ts, m)]
    extra = []
    for s in input_sets[1:]:
        acc = [0] * n
        for inp in s:
            fi = [(beta + v) % R_MOD for v in inp]
            batch_invert(fi)
            acc = [(a + b) % R_MOD for a, b in zip(acc, fi)]
        extra.append(acc)
    u = n - (bf + 1)
    last_z = 0
    zs = []
    for grand in [base] + extra:
        z, state = [], 0
        for v in [last_z] + grand:
            state = (state + v) % R_MOD
            z.append(state)
            if len(z) == n - bf:
                break
        last_z = z[u]
        zs.append(z)
    return zs


def blind_to_n(z: List[int], n: int, rng: random.Random) -> List[int]:
    """plonk/prover.rs: the raw z vectors are extended to n rows with random blinding values."""
    return list(z) + [rng.randrange(R_MOD) for _ in range(n - len(z))]


def shuffle_commit_product(cs, domain, group, theta, beta, advice, fixed, instance) -> List[int]:
    """shuffle/prover.rs:60-157 -> z (length n - blinding_factors)"""
    n = domain.n
    bf = cs.blinding_factors()
    comp = lambda ex: evaluate_with_theta(ex, n, 1, fixed, advice, instance, theta)  # noqa: E731
    inputs = [comp(a["input_expressions"]) for a in group]
    shuffles = [comp(a["shuffle_expressions"]) for a in group]
    challenges = [pow(beta, 1 + i, R_MOD) for i in range(len(group))]
    prod = [1] * n
    for ex, ch in zip(shuffles, challenges):
        prod = [p * ((ch + v) % R_MOD) % R_MOD for p, v in zip(prod, ex)]
    batch_invert(prod)
    for ex, ch in zip(inputs, challenges):
        prod = [p * ((ch + v) % R_MOD) % R_MOD for p, v in zip(prod, ex)]
    z, state = [], 1
    for v in [1] + prod:
        state = state * v % R_MOD
        z.append(state)
        if len(z) == n - bf:
            break
    return z


# --------------------------------------------------------------------------
# evaluate_h (non-cuda variant), evaluation.rs:778-1226
# --------------------------------------------------------------------------
def evaluate_h(ev: Evaluator, cs: ConstraintSystem, domain: EvaluationDomain, fixed_cosets, advice_cosets,
               instance_cosets, l0, l_last, l_active_row, sigma_cosets, y, beta, gamma, theta,
               lookups, shuffles, permutation_sets, zeta: Optional[int] = None,
               values: Optional[List[int]] = None) -> List[int]:
    """lookups: [{"z_cosets": [...], "m_coset": [...]}], shuffles: [product_coset], permutation_sets:
    [z coset per set]; all cosets are extended-domain evaluation lists.  One circuit instance: the reference loops
    over `advice.iter()` = the instances of a batch (evaluation.rs:839-845) with ONE accumulator; `values` is that
    accumulator as the previous instance left it (None = zeros, the first instance)."""
    size = domain.extended_len()
    rot_scale = 1 << (domain.extended_k - domain.k)
    ext_omega = domain.extended_omega
    zeta = domain.g_coset if zeta is None else zeta
    fixed, advice, instance = fixed_cosets, advice_cosets, instance_cosets
    values = [0] * size if values is None else list(values)
    n_lookups = len(cs.lookups)
    table_values = [[0] * size for _ in range(n_lookups)]
    input_product = [[0] * size for _ in range(n_lookups)]
    input_product_sum = [[0] * size for _ in range(n_lookups)]
    extra_product: List[List[int]] = []
    extra_sum: List[List[int]] = []
    for lk in ev.lookup_results:
        for _ in range(len(lk[1]) - 1):
            extra_product.append([0] * size)
            extra_sum.append([0] * size)
    shuffle_in = [[0] * size for _ in ev.shuffle_results]
    shuffle_tab = [[0] * size for _ in ev.shuffle_results]

    # expressions :846-1001
    for idx in range(size):
        rotations = [(idx + rot * rot_scale) % size for rot in ev.rotations]
        inter = [0] * len(ev.calculations)
        ce = lambda c: calc_evaluate(c, rotations, ev.constants, inter, fixed, advice, instance, beta, gamma, theta)  # noqa: E731
        for i, calc in enumerate(ev.calculations):
            inter[i] = ce(calc)
        v = values[idx]
        for part in ev.value_parts:
            v = (v * y + _get(part, rotations, ev.constants, inter, fixed, advice, instance)) % R_MOD
        values[idx] = v
        off = 0
        for t, res in enumerate(ev.lookup_results):
            table_values[t][idx] = ce(res[0])
            input_product[t][idx] = ce(res[1][0])
            input_product_sum[t][idx] = ce(res[2][0])
            for i in range(1, len(res[1])):
                extra_product[off][idx] = ce(res[1][i])
                extra_sum[off][idx] = ce(res[2][i])
                off += 1
        for i, res in enumerate(ev.shuffle_results):
            shuffle_in[i][idx] = ce(res[0])
            shuffle_tab[i][idx] = ce(res[1])

    bf = cs.blinding_factors()
    last_rotation = -(bf + 1)
    # permutations :1005-1090
    sets = permutation_sets
    if sets:
        chunk_len = cs.degree() - 2
        delta_start = beta * zeta % R_MOD
        first_set, last_set = sets[0], sets[-1]
        beta_term = 1
        for idx in range(size):
            r_next = (idx + rot_scale) % size
            r_last = (idx + last_rotation * rot_scale) % size
            v = values[idx]
            v = (v * y + (1 - first_set[idx]) * l0[idx]) % R_MOD
            v = (v * y + (last_set[idx] * last_set[idx] - last_set[idx]) * l_last[idx]) % R_MOD
            for si in range(1, len(sets)):
                v = (v * y + (sets[si][idx] - sets[si - 1][r_last]) * l0[idx]) % R_MOD
            current_delta = delta_start * beta_term % R_MOD
            for si, z in enumerate(sets):
                columns = cs.permutation_columns[si * chunk_len:(si + 1) * chunk_len]
                cosets = sigma_cosets[si * chunk_len:(si + 1) * chunk_len]
                left = z[r_next]
                for column, perm in zip(columns, cosets):
                    vals = _column(column, fixed, advice, instance)
                    left = left * ((vals[idx] + beta * perm[idx] + gamma) % R_MOD) % R_MOD
                right = z[idx]
                for column in columns:
                    vals = _column(column, fixed, advice, instance)
                    right = right * ((vals[idx] + current_delta + gamma) % R_MOD) % R_MOD
                    current_delta = current_delta * FR_DELTA % R_MOD
                v = (v * y + (left - right) * l_active_row[idx]) % R_MOD
            values[idx] = v
            beta_term = beta_term * ext_omega % R_MOD

    # lookups :1104-1180
    off = 0
    for li, lookup in enumerate(lookups):
        z_sets = lookup["z_cosets"]
        m_coset = lookup["m_coset"]
        sets_len = len(z_sets)
        ext_prod = extra_product[off:off + sets_len - 1]
        ext_sum = extra_sum[off:off + sets_len - 1]
        off += sets_len - 1
        table, ip, ips = table_values[li], input_product[li], input_product_sum[li]
        for idx in range(size):
            r_next = (idx + rot_scale) % size
            r_last = (idx + last_rotation * rot_scale) % size
            v = values[idx]
            v = (v * y + z_sets[0][idx] * l0[idx]) % R_MOD
            v = (v * y + z_sets[sets_len - 1][idx] * l_last[idx]) % R_MOD
            dz = (z_sets[0][r_next] - z_sets[0][idx]) % R_MOD
            v = (v * y + ((dz * table[idx] + m_coset[idx]) * ip[idx] - table[idx] * ips[idx]) * l_active_row[idx]) % R_MOD
            for i in range(1, sets_len):
                v = (v * y + (z_sets[i][idx] - z_sets[i - 1][r_last]) * l0[idx]) % R_MOD
            for i in range(1, sets_len):
                dz = (z_sets[i][r_next] - z_sets[i][idx]) % R_MOD
                v = (v * y + (dz * ext_prod[i - 1][idx] - ext_sum[i - 1][idx]) * l_active_row[idx]) % R_MOD
            values[idx] = v

    # shuffles :1184-1220
    for si, product_coset in enumerate(shuffles):
        for idx in range(size):
            r_next = (idx + rot_scale) % size
            v = values[idx]
            v = (v * y + (1 - product_coset[idx]) * l0[idx]) % R_MOD
            v = (v * y + (product_coset[idx] * product_coset[idx] - product_coset[idx]) * l_last[idx]) % R_MOD
            v = (v * y + (product_coset[r_next] * shuffle_tab[si][idx] - product_coset[idx] * shuffle_in[si][idx])
                 * l_active_row[idx]) % R_MOD
            values[idx] = v
    return values
