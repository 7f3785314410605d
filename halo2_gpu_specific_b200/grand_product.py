"""Device-resident construction of the prover's z columns -- the step right before each
commit_lagrange_and_ifft:

  permutation_commit      plonk/permutation/prover.rs:47-165   (grand product per column chunk)
  logup_commit_z          plonk/logup/prover.rs:263-336        (grand sum per input set)
  shuffle_commit_product  plonk/shuffle/prover.rs:107-141      (grand product)

Each is: per-row expressions (the quotient engine's interpreter kernel at rows = n, rot_scale = 1;
evaluate_with_theta, plonk/evaluation.rs:2330-2398), batch_invert (arithmetic.rs:840-844), one more
per-row pass, and a prefix product / sum whose start value is read from the previous set's z on the
device (last_z, permutation/prover.rs:146,160).  Columns are uploaded once; only the finished z columns
come back.  The random blinding rows are supplied by the caller (the reference draws them from its RNG).
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error, check, lib, require_gpu
from .evaluation import CH_BETA, CH_GAMMA, CH_THETA, DELTA, DeviceBuffer, QuotientProgram

R = _fr.R_MOD


class ExprCompiler:
    """Expression -> Calculation list, the scheme of Evaluator::add_expression (plonk/evaluation.rs:671-776)
    without the cross-expression cache.  Expressions are the reference's enum as tuples:
    ("Constant", v) | ("Fixed"|"Advice"|"Instance", column, rotation) | ("Negated", e) | ("Sum", a, b)
    | ("Product", a, b) | ("Scaled", e, v)."""

    def __init__(self):
        self.rotations: List[int] = [0]
        self.constants: List[int] = [0, 1]
        self.calcs: List[tuple] = []

    def emit(self, c: tuple):
        self.calcs.append(c)
        return ("Intermediate", len(self.calcs) - 1)

    def const(self, v: int):
        v %= R
        if v not in self.constants:
            self.constants.append(v)
        return ("Constant", self.constants.index(v))

    def rot(self, r: int) -> int:
        if r not in self.rotations:
            self.rotations.append(r)
        return self.rotations.index(r)

    def expr(self, e):
        t = e[0]
        if t == "Constant":
            return self.const(e[1])
        if t in ("Fixed", "Advice", "Instance"):
            return (t, e[1], self.rot(e[2]))
        if t == "Negated":
            return self.emit(("Negate", self.expr(e[1])))
        if t == "Sum":
            if e[2][0] == "Negated":
                return self.emit(("Sub", self.expr(e[1]), self.expr(e[2][1])))
            return self.emit(("Add", self.expr(e[1]), self.expr(e[2])))
        if t == "Product":
            return self.emit(("Mul", self.expr(e[1]), self.expr(e[2])))
        if t == "Scaled":
            return self.emit(("Mul", self.expr(e[1]), self.const(e[2])))
        raise B2Error(B2_ERR_ARG, f"unknown Expression {e!r}")

    def compress(self, expressions):
        """evaluate_with_theta: fold with theta (evaluation.rs:2387-2393)"""
        acc = None
        for e in expressions:
            v = self.expr(e)
            acc = v if acc is None else self.emit(("LcTheta", acc, v))
        return acc


def _scan(op: int, d_in: int, n_in: int, d_out: int, n_out: int, init: int | None = None, d_init: int = 0) -> None:
    init_arr = _fr.to_mont(init) if init is not None else None
    check(lib().b2_prefix_scan_dev(op, ctypes.c_void_p(d_in), n_in,
                                   ctypes.c_void_p(init_arr.ctypes.data) if init_arr is not None else None,
                                   ctypes.c_void_p(d_init) if d_init else None, ctypes.c_void_p(d_out), n_out, None))


def _invert(d_a: int, n: int) -> None:
    check(lib().b2_batch_invert_dev(ctypes.c_void_p(d_a), n, None))


class _Columns:
    """fixed / advice / instance Lagrange columns resident on the device"""

    def __init__(self, fixed, advice, instance, n: int):
        self.n = n
        self.counts = (len(fixed), len(advice), len(instance))
        cols = [np.asarray(c, dtype=np.uint64).reshape(n, 4) for c in list(fixed) + list(advice) + list(instance)]
        self.buf = DeviceBuffer(max(1, len(cols)) * n)
        if cols:
            self.buf.upload(np.stack(cols))
        p = [self.buf.ptr + i * n * 32 for i in range(len(cols))]
        nf, na, _ = self.counts
        self.fixed, self.advice, self.instance = p[:nf], p[nf:nf + na], p[nf + na:]

    @classmethod
    def resident(cls, fixed_ptrs, advice_ptrs, instance_ptrs, n: int) -> "_Columns":
        """columns that already live on the device (device pointers); nothing is owned"""
        self = cls.__new__(cls)
        self.n = n
        self.fixed, self.advice, self.instance = list(fixed_ptrs), list(advice_ptrs), list(instance_ptrs)
        self.counts = (len(self.fixed), len(self.advice), len(self.instance))
        self.buf = None
        return self

    def of(self, kind: str):
        return {"Fixed": self.fixed, "Advice": self.advice, "Instance": self.instance}[kind]

    def free(self):
        if self.buf is not None:
            self.buf.free()


def _run(comp: ExprCompiler, result, cols: _Columns, aux: Sequence[int], challenges: Sequence[int], out_ptr: int,
         k: int, x0=None, x_step=None) -> None:
    nf, na, ni = cols.counts
    prog = QuotientProgram(comp.rotations, comp.constants, comp.calcs, result, nf, na, ni, len(aux), len(challenges))
    try:
        prog.eval(k, 1, cols.fixed, cols.advice, cols.instance, list(aux), list(challenges), out_ptr, x0=x0, x_step=x_step)
    finally:
        prog.free()


def _h2d(dst_ptr: int, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    check(lib().b2_memcpy_h2d(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(a.ctypes.data), a.nbytes))


def permutation_commit_dev(domain, permutation_columns, degree: int, blinding_factors: int, sig_ptrs: Sequence[int],
                           cols: _Columns, beta: int, gamma: int, blinds: Sequence[np.ndarray], z_ptrs: Sequence[int]) -> None:
    """permutation/prover.rs:47-165 on resident columns: sig_ptrs are the permutation polynomials in Lagrange form
    (pkey.permutations), z_ptrs[s] receives the finished z column of set s (n rows, blinding rows included)."""
    n, k = domain.n, domain.k
    chunk_len = degree - 2                                            # :66
    m = len(permutation_columns)
    n_sets = (m + chunk_len - 1) // chunk_len
    if len(blinds) != n_sets or len(z_ptrs) != n_sets or len(sig_ptrs) != m:
        raise B2Error(B2_ERR_ARG, f"expected {n_sets} sets of blinding rows / outputs and {m} sigma columns")
    work = DeviceBuffer(max(1, n_sets) * n)     # denominators, then the fractions
    try:
        # challenges: beta, gamma, theta (unused), then beta * DELTA^j per column (delta_omega * beta, :108-121)
        ch = [beta % R, gamma % R, 0]
        d = beta % R
        for _ in range(m):
            ch.append(d)
            d = d * DELTA % R
        for s in range(n_sets):
            chunk = list(enumerate(permutation_columns))[s * chunk_len:(s + 1) * chunk_len]
            # denominators: prod_j (beta * sigma_j + gamma + v_j)   :90-101
            c = ExprCompiler()
            acc = None
            for j, (kind, idx) in chunk:
                t = c.emit(("Mul", ("Challenge", CH_BETA), ("Aux", j, 0)))
                t = c.emit(("AddChallenge", c.emit(("Add", t, (kind, idx, 0))), "Gamma"))
                acc = t if acc is None else c.emit(("Mul", acc, t))
            _run(c, acc, cols, list(sig_ptrs), ch, work.ptr + s * n * 32, k)
        _invert(work.ptr, n_sets * n)                                  # :104 (all sets in one launch)
        for s in range(n_sets):
            chunk = list(enumerate(permutation_columns))[s * chunk_len:(s + 1) * chunk_len]
            # fractions: inverse * prod_j (delta^j * omega^i * beta + gamma + v_j)   :108-121
            c = ExprCompiler()
            acc = ("Aux", 0, 0)
            for j, (kind, idx) in chunk:
                t = c.emit(("Mul", ("CosetX",), ("Challenge", 3 + j)))
                t = c.emit(("AddChallenge", c.emit(("Add", t, (kind, idx, 0))), "Gamma"))
                acc = c.emit(("Mul", acc, t))
            frac = work.ptr + s * n * 32
            _run(c, acc, cols, [frac], ch, frac, k, x0=1, x_step=domain._omega)
            # z[0] = last_z, z[i + 1] = z[i] * fraction[i]   :135-152
            if s == 0:
                _scan(0, frac, n, z_ptrs[s], n, init=1)
            else:
                last = z_ptrs[s - 1] + (n - (blinding_factors + 1)) * 32                   # :160
                _scan(0, frac, n, z_ptrs[s], n, d_init=last)
            _h2d(z_ptrs[s] + (n - blinding_factors) * 32,
                 np.asarray(blinds[s], dtype=np.uint64).reshape(blinding_factors, 4))       # :156-158
    finally:
        work.free()


def permutation_commit(domain, permutation_columns, degree: int, blinding_factors: int, sigmas, advice, fixed, instance,
                       beta: int, gamma: int, blinds: Sequence[np.ndarray]) -> List[np.ndarray]:
    """permutation/prover.rs:47-165.  sigmas: the permutation polynomials in Lagrange form
    (pkey.permutations); blinds[s]: (blinding_factors, 4) values for the last rows of set s.
    Returns the z columns (Lagrange basis, (n, 4) Montgomery), one per column chunk."""
    require_gpu()
    n = domain.n
    chunk_len = degree - 2
    m = len(permutation_columns)
    n_sets = (m + chunk_len - 1) // chunk_len
    if len(blinds) != n_sets:
        raise B2Error(B2_ERR_ARG, f"expected blinding rows for {n_sets} sets")
    cols = _Columns(fixed, advice, instance, n)
    sig = DeviceBuffer(max(1, m) * n)
    zbuf = DeviceBuffer(max(1, n_sets) * n)
    try:
        if m:
            sig.upload(np.stack([np.asarray(s, dtype=np.uint64).reshape(n, 4) for s in sigmas]))
        permutation_commit_dev(domain, permutation_columns, degree, blinding_factors,
                               [sig.ptr + j * n * 32 for j in range(m)], cols, beta, gamma, blinds,
                               [zbuf.ptr + s * n * 32 for s in range(n_sets)])
        out = zbuf.download(n_sets * n).reshape(n_sets, n, 4) if n_sets else np.zeros((0, n, 4), np.uint64)
        return [out[s] for s in range(n_sets)]
    finally:
        cols.free(); sig.free(); zbuf.free()


def logup_commit_z_dev(domain, lookup, blinding_factors: int, cols: _Columns, m_ptr: int, theta: int, beta: int,
                       z_ptrs: Sequence[int]) -> None:
    """logup/prover.rs:263-336 (with the compression of :83-112 fused in) on resident columns; m_ptr: m(X) values;
    z_ptrs[i] receives the first n - blinding_factors rows of the z column of input set i."""
    n, k = domain.n, domain.k
    sets = lookup["input_expressions_sets"]
    if len(z_ptrs) != len(sets):
        raise B2Error(B2_ERR_ARG, f"expected {len(sets)} output columns")
    n_inputs = sum(len(s) for s in sets)
    inv = DeviceBuffer((n_inputs + 1) * n)       # beta + compressed input, per input; then beta + table
    grand = DeviceBuffer(len(sets) * n)
    ch = [beta % R, 0, theta % R]
    try:
        slot = 0
        for s in sets:                                                 # :277-283, :308-316
            for inp in s:
                c = ExprCompiler()
                _run(c, c.emit(("AddChallenge", c.compress(inp), "Beta")), cols, [], ch, inv.ptr + slot * n * 32, k)
                slot += 1
        c = ExprCompiler()                                             # :288-296
        _run(c, c.emit(("AddChallenge", c.compress(lookup["table_expressions"]), "Beta")), cols, [], ch,
             inv.ptr + slot * n * 32, k)
        _invert(inv.ptr, (n_inputs + 1) * n)
        inv_ptrs = [inv.ptr + i * n * 32 for i in range(n_inputs + 1)]
        slot = 0
        for si, s in enumerate(sets):
            c = ExprCompiler()
            acc = None
            for _ in s:
                v = ("Aux", slot, 0)
                acc = v if acc is None else c.emit(("Add", acc, v))
                slot += 1
            if si == 0:                                                # :297-305: sum - table_inv * m
                acc = c.emit(("Sub", acc, c.emit(("Mul", ("Aux", n_inputs, 0), ("Aux", n_inputs + 1, 0)))))
            _run(c, acc, cols, inv_ptrs + [m_ptr], ch, grand.ptr + si * n * 32, k)
        u = n - (blinding_factors + 1)
        n_out = n - blinding_factors
        for si in range(len(sets)):                                    # :318-336
            if si == 0:
                _scan(1, grand.ptr, n, z_ptrs[si], n_out, init=0)
            else:
                _scan(1, grand.ptr + si * n * 32, n, z_ptrs[si], n_out, d_init=z_ptrs[si - 1] + u * 32)
    finally:
        inv.free(); grand.free()


def logup_commit_z(domain, lookup, blinding_factors: int, advice, fixed, instance, multiplicity, theta: int, beta: int
                   ) -> List[np.ndarray]:
    """logup/prover.rs:263-336 with the compression of :83-112 fused in.  lookup: {"table_expressions",
    "input_expressions_sets"} (Expression tuples); multiplicity: m(X) values (n, 4).
    Returns the raw z vectors (n - blinding_factors rows each), one per input set."""
    require_gpu()
    n = domain.n
    sets = lookup["input_expressions_sets"]
    cols = _Columns(fixed, advice, instance, n)
    zbuf = DeviceBuffer(len(sets) * n)
    mbuf = DeviceBuffer(n).upload(np.asarray(multiplicity, dtype=np.uint64).reshape(n, 4))
    try:
        logup_commit_z_dev(domain, lookup, blinding_factors, cols, mbuf.ptr, theta, beta,
                           [zbuf.ptr + i * n * 32 for i in range(len(sets))])
        n_out = n - blinding_factors
        out = zbuf.download(len(sets) * n).reshape(len(sets), n, 4)
        return [out[i, :n_out].copy() for i in range(len(sets))]
    finally:
        cols.free(); zbuf.free(); mbuf.free()


def shuffle_commit_product_dev(domain, group, blinding_factors: int, cols: _Columns, theta: int, beta: int, z_ptr: int
                               ) -> None:
    """shuffle/prover.rs:60-141 on resident columns; z_ptr receives the first n - blinding_factors rows of z."""
    n, k = domain.n, domain.k
    work = DeviceBuffer(n)
    ch = [beta % R, 0, theta % R]
    try:
        def product(which: str, c: ExprCompiler, acc):
            for i, arg in enumerate(group):                           # challenges beta^(1 + i), :114-116
                acc = c.emit(("LcChallenge", c.compress(arg[which]), acc, "Beta", i + 1))
            return acc
        c = ExprCompiler()
        _run(c, product("shuffle_expressions", c, c.const(1)), cols, [], ch, work.ptr, k)   # :118-129
        _invert(work.ptr, n)                                                                  # :132
        c = ExprCompiler()
        _run(c, product("input_expressions", c, ("Aux", 0, 0)), cols, [work.ptr], ch, work.ptr, k)   # :134
        _scan(0, work.ptr, n, z_ptr, n - blinding_factors, init=1)                            # :137-146
    finally:
        work.free()


def shuffle_commit_product(domain, group, blinding_factors: int, advice, fixed, instance, theta: int, beta: int
                           ) -> np.ndarray:
    """shuffle/prover.rs:60-141.  group: [{"input_expressions", "shuffle_expressions"}].
    Returns z (n - blinding_factors rows)."""
    require_gpu()
    n = domain.n
    cols = _Columns(fixed, advice, instance, n)
    zbuf = DeviceBuffer(n)
    try:
        shuffle_commit_product_dev(domain, group, blinding_factors, cols, theta, beta, zbuf.ptr)
        return zbuf.download(n - blinding_factors)
    finally:
        cols.free(); zbuf.free()


def compress_expressions_dev(domain, expression_lists, cols: _Columns, theta: int, out_ptr: int) -> None:
    """evaluate_with_theta (plonk/evaluation.rs:2330-2398) per expression list on resident columns; list i goes to
    out_ptr + i * n * 32"""
    n, k = domain.n, domain.k
    for i, exprs in enumerate(expression_lists):
        c = ExprCompiler()
        _run(c, c.emit(("Store", c.compress(exprs))), cols, [], [0, 0, theta % R], out_ptr + i * n * 32, k)


def compress_expressions(domain, expression_lists, advice, fixed, instance, theta: int) -> np.ndarray:
    """evaluate_with_theta (plonk/evaluation.rs:2330-2398) for several expression lists over the same Lagrange
    columns: what logup's `compress` computes before the multiplicities are counted (logup/prover.rs:83-112).
    Returns (len(expression_lists), n, 4)."""
    require_gpu()
    n = domain.n
    m = len(expression_lists)
    cols = _Columns(fixed, advice, instance, n)
    out = DeviceBuffer(max(1, m) * n)
    try:
        compress_expressions_dev(domain, expression_lists, cols, theta, out.ptr)
        return out.download(m * n).reshape(m, n, 4) if m else np.zeros((0, n, 4), np.uint64)
    finally:
        cols.free(); out.free()


def batch_invert(a: np.ndarray) -> np.ndarray:
    """arithmetic.rs:840-844, in place on a host array"""
    require_gpu()
    if not a.flags.c_contiguous or a.dtype != np.uint64:
        raise B2Error(B2_ERR_ARG, "expected a C-contiguous uint64 array")
    check(lib().b2_batch_invert(ctypes.c_void_p(a.ctypes.data), a.size // 4))
    return a


def prefix_scan(op: str, a: np.ndarray, init: int | None = None, n_out: int | None = None) -> np.ndarray:
    """out[0] = init, out[i + 1] = out[i] (op) a[i]; op: "product" (mul_acc, arithmetic.rs:806-836) or "sum"."""
    require_gpu()
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    n_out = a.shape[0] + 1 if n_out is None else n_out
    out = np.empty((n_out, 4), dtype=np.uint64)
    init_arr = _fr.to_mont(init) if init is not None else None
    check(lib().b2_prefix_scan({"product": 0, "sum": 1}[op], ctypes.c_void_p(a.ctypes.data), a.shape[0],
                               ctypes.c_void_p(init_arr.ctypes.data) if init_arr is not None else None,
                               ctypes.c_void_p(out.ctypes.data), n_out))
    return out
