"""Mirror of plonk::Evaluator::evaluate_h (halo2_proofs/src/plonk/evaluation.rs:778-1226) over the
C ABI (b2_quotient_*): the whole numerator of the quotient polynomial -- custom gates, permutation,
logup lookups and shuffles, folded with y -- is ONE program evaluated by one kernel per proof, over
cosets that never leave the GPU (coefficient form in, extended evaluations or h(X) coefficients out).

Data interchange is the reference's own: `rotations`, `constants`, `calculations`, `value_parts`,
`lookup_results`, `shuffle_results` are the fields of Evaluator (evaluation.rs:270-298) with the enums
written as tuples:
  ValueSource  ("Constant", i) | ("Intermediate", i) | ("Fixed"|"Advice"|"Instance", column, rotation_index)
  Calculation  ("Add"|"Sub"|"Mul", a, b) | ("Negate", a) | ("LcChallenge", a, b, "Beta"|"Gamma", p)
               | ("LcTheta", a, b) | ("AddChallenge", a, "Beta"|"Gamma") | ("Store", a)
Field elements cross this module as canonical Python ints (constants, challenges) or as (n, 4) uint64
Montgomery arrays (polynomials), the layout of Vec<Fr>.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _fr
from ._lib import B2_ERR_ARG, B2Error, NttDesc, check, lib, require_gpu

# include/b2pcs.h
Q_CONSTANT, Q_INTERMEDIATE, Q_FIXED, Q_ADVICE, Q_INSTANCE, Q_AUX, Q_CHALLENGE, Q_COSET_X = range(8)
OP_ADD, OP_SUB, OP_MUL, OP_NEGATE, OP_LC_CHALLENGE, OP_MUL_CH_ADD, OP_ADD_CHALLENGE, OP_STORE = range(8)
# challenge table layout used by the programs built here
CH_BETA, CH_GAMMA, CH_THETA, CH_Y, CH_FIRST_DELTA = 0, 1, 2, 3, 4

DELTA = pow(_fr.GENERATOR, 1 << _fr.S, _fr.R_MOD)   # Fr::DELTA


class QSrc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("index", ctypes.c_uint32), ("rotation", ctypes.c_uint32)]


class QCalc(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint32), ("a", QSrc), ("b", QSrc), ("challenge", ctypes.c_uint32),
                ("power", ctypes.c_uint32)]


class QProgramDesc(ctypes.Structure):
    _fields_ = [("rotations", ctypes.c_void_p), ("n_rotations", ctypes.c_uint32),
                ("constants", ctypes.c_void_p), ("n_constants", ctypes.c_uint32),
                ("calcs", ctypes.c_void_p), ("n_calcs", ctypes.c_uint32),
                ("result", QSrc),
                ("n_fixed", ctypes.c_uint32), ("n_advice", ctypes.c_uint32), ("n_instance", ctypes.c_uint32),
                ("n_aux", ctypes.c_uint32), ("n_challenges", ctypes.c_uint32)]


class QArgs(ctypes.Structure):
    _fields_ = [("log_rows", ctypes.c_uint32), ("rot_scale", ctypes.c_uint32),
                ("fixed", ctypes.c_void_p), ("advice", ctypes.c_void_p), ("instance", ctypes.c_void_p),
                ("aux", ctypes.c_void_p), ("challenges", ctypes.c_void_p),
                ("x0", ctypes.c_void_p), ("x_step", ctypes.c_void_p),
                ("scale", ctypes.c_void_p), ("scale_len", ctypes.c_uint32),
                ("out", ctypes.c_void_p), ("out_stride", ctypes.c_uint64), ("out_offset", ctypes.c_uint64),
                ("stream", ctypes.c_void_p), ("row_begin", ctypes.c_uint64), ("row_count", ctypes.c_uint64)]


_KIND = {"Constant": Q_CONSTANT, "Intermediate": Q_INTERMEDIATE, "Fixed": Q_FIXED, "Advice": Q_ADVICE,
         "Instance": Q_INSTANCE, "Aux": Q_AUX, "Challenge": Q_CHALLENGE, "CosetX": Q_COSET_X}
_CH = {"Beta": CH_BETA, "Gamma": CH_GAMMA, "Theta": CH_THETA, "Y": CH_Y}


def _src(s) -> QSrc:
    kind = _KIND.get(s[0])
    if kind is None:
        raise B2Error(B2_ERR_ARG, f"unknown ValueSource {s!r}")
    index = s[1] if len(s) > 1 else 0
    rot = s[2] if len(s) > 2 else 0
    return QSrc(kind, index, rot)


def _calc(c) -> QCalc:
    t = c[0]
    z = QSrc(0, 0, 0)
    if t == "Add":
        return QCalc(OP_ADD, _src(c[1]), _src(c[2]), 0, 0)
    if t == "Sub":
        return QCalc(OP_SUB, _src(c[1]), _src(c[2]), 0, 0)
    if t == "Mul":
        return QCalc(OP_MUL, _src(c[1]), _src(c[2]), 0, 0)
    if t == "Negate":
        return QCalc(OP_NEGATE, _src(c[1]), z, 0, 0)
    if t == "LcChallenge":
        return QCalc(OP_LC_CHALLENGE, _src(c[1]), _src(c[2]), _CH[c[3]], c[4])
    if t == "LcTheta":
        return QCalc(OP_MUL_CH_ADD, _src(c[1]), _src(c[2]), CH_THETA, 0)
    if t == "MulChAdd":   # a * challenges[c[3]] + b (the y fold)
        return QCalc(OP_MUL_CH_ADD, _src(c[1]), _src(c[2]), c[3], 0)
    if t == "AddChallenge":
        return QCalc(OP_ADD_CHALLENGE, _src(c[1]), z, _CH[c[2]], 0)
    if t == "Store":
        return QCalc(OP_STORE, _src(c[1]), z, 0, 0)
    raise B2Error(B2_ERR_ARG, f"unknown Calculation {c!r}")


class QuotientProgram:
    """Handle of a lowered program (b2_quotient_program_create).  Creation is host-only."""

    def __init__(self, rotations: Sequence[int], constants: Sequence[int], calcs: Sequence[tuple], result,
                 n_fixed: int, n_advice: int, n_instance: int, n_aux: int, n_challenges: int):
        rot = np.asarray(list(rotations), dtype=np.int32)
        cst = np.stack([_fr.to_mont(c) for c in constants]) if len(constants) else np.zeros((0, 4), np.uint64)
        arr = (QCalc * max(1, len(calcs)))(*[_calc(c) for c in calcs])
        d = QProgramDesc()
        d.rotations, d.n_rotations = rot.ctypes.data, len(rot)
        d.constants, d.n_constants = cst.ctypes.data, len(cst)
        d.calcs, d.n_calcs = ctypes.addressof(arr), len(calcs)
        d.result = _src(result)
        d.n_fixed, d.n_advice, d.n_instance, d.n_aux, d.n_challenges = n_fixed, n_advice, n_instance, n_aux, n_challenges
        h = ctypes.c_uint64()
        check(lib().b2_quotient_program_create(ctypes.byref(d), ctypes.byref(h)))
        self.handle = h.value
        self.n_fixed, self.n_advice, self.n_instance, self.n_aux, self.n_challenges = \
            n_fixed, n_advice, n_instance, n_aux, n_challenges

    def info(self):
        v = [ctypes.c_uint32() for _ in range(4)]
        check(lib().b2_quotient_program_info(ctypes.c_uint64(self.handle), *[ctypes.byref(x) for x in v]))
        out = dict(zip(("n_instr", "n_slots", "n_mul", "n_addsub"), (x.value for x in v)))
        sh, gl = ctypes.c_uint32(), ctypes.c_uint32()
        check(lib().b2_quotient_program_slot_classes(ctypes.c_uint64(self.handle), ctypes.byref(sh), ctypes.byref(gl)))
        out["n_slots_shared"], out["n_slots_global"] = sh.value, gl.value
        return out

    def dump(self):
        """(instructions, result word, derived (challenge, power) pairs); an instruction is (op, dst, a_word, b_word),
        or (op, dst, a, b, c, d) for the fused a * b +- c * d (op 5 / 6: its extension entry is folded in here)"""
        n = self.info()["n_instr"]
        words = np.zeros(max(1, n) * 4, dtype=np.uint32)
        der = np.zeros(2 * 4096, dtype=np.uint32)
        res, nd = ctypes.c_uint32(), ctypes.c_uint32()
        check(lib().b2_quotient_program_dump(ctypes.c_uint64(self.handle), ctypes.c_void_p(words.ctypes.data), words.size,
                                             ctypes.byref(res), ctypes.c_void_p(der.ctypes.data), der.size,
                                             ctypes.byref(nd)))
        instr, i = [], 0
        while i < n:
            op, dst = int(words[4 * i]) & 0xff, int(words[4 * i]) >> 8
            if op in (5, 6):
                assert i + 1 < n and int(words[4 * i + 4]) & 0xff == 7, "fused instruction without its extension entry"
                instr.append((op, dst, int(words[4 * i + 1]), int(words[4 * i + 2]), int(words[4 * i + 3]),
                              int(words[4 * i + 5])))
                i += 2
            else:
                instr.append((op, dst, int(words[4 * i + 1]), int(words[4 * i + 2])))
                i += 1
        return instr, res.value, [(int(der[2 * i]), int(der[2 * i + 1])) for i in range(nd.value)]

    def free(self):
        if self.handle:
            lib().b2_quotient_program_free(ctypes.c_uint64(self.handle))
            self.handle = 0

    def eval(self, log_rows: int, rot_scale: int, fixed, advice, instance, aux, challenges: Sequence[int], out_ptr: int,
             x0: Optional[int] = None, x_step: Optional[int] = None, scale: Optional[np.ndarray] = None,
             out_stride: int = 1, out_offset: int = 0, stream: int = 0, row_begin: int = 0, row_count: int = 0) -> None:
        """fixed / advice / instance / aux: sequences of DEVICE pointers (ints)."""
        require_gpu()

        def table(ptrs, n):
            if len(ptrs) != n:
                raise B2Error(B2_ERR_ARG, f"expected {n} column pointers, got {len(ptrs)}")
            return (ctypes.c_void_p * max(1, n))(*[ctypes.c_void_p(int(p)) for p in ptrs])

        tf, ta, ti, tx = (table(fixed, self.n_fixed), table(advice, self.n_advice),
                          table(instance, self.n_instance), table(aux, self.n_aux))
        if len(challenges) != self.n_challenges:
            raise B2Error(B2_ERR_ARG, f"expected {self.n_challenges} challenges, got {len(challenges)}")
        ch = np.stack([_fr.to_mont(c) for c in challenges]) if len(challenges) else np.zeros((1, 4), np.uint64)
        a = QArgs()
        a.log_rows, a.rot_scale = log_rows, rot_scale
        a.fixed, a.advice = ctypes.addressof(tf), ctypes.addressof(ta)
        a.instance, a.aux = ctypes.addressof(ti), ctypes.addressof(tx)
        a.challenges = ch.ctypes.data
        keep = []
        if x0 is not None:
            mx0, mstep = _fr.to_mont(x0), _fr.to_mont(x_step)
            keep += [mx0, mstep]
            a.x0, a.x_step = mx0.ctypes.data, mstep.ctypes.data
        if scale is not None:
            scale = np.ascontiguousarray(scale, dtype=np.uint64).reshape(-1, 4)
            a.scale, a.scale_len = scale.ctypes.data, scale.shape[0]
        a.out, a.out_stride, a.out_offset, a.stream = out_ptr, out_stride, out_offset, stream
        a.row_begin, a.row_count = row_begin, row_count
        check(lib().b2_quotient_eval(ctypes.c_uint64(self.handle), ctypes.byref(a)))


class BufferPool:
    """Exact-size free lists of device allocations.  A proof allocates the same few dozen sizes every time (columns,
    blocks of columns, scratch of the z-column kernels); cudaMalloc / cudaFree of 128 MiB .. 8 GiB blocks cost
    milliseconds each and synchronise the device, so an engine that proves repeatedly keeps what it freed."""

    def __init__(self):
        self._free = {}
        self.cached_bytes = 0

    def take(self, nbytes: int):
        lst = self._free.get(nbytes)
        if lst:
            self.cached_bytes -= nbytes
            return lst.pop()
        return None

    def give(self, ptr: int, nbytes: int) -> None:
        self._free.setdefault(nbytes, []).append(ptr)
        self.cached_bytes += nbytes

    def trim(self) -> None:
        for lst in self._free.values():
            for p in lst:
                lib().b2_dev_free(ctypes.c_void_p(p))
        self._free = {}
        self.cached_bytes = 0


_ACTIVE_POOL: Optional[BufferPool] = None


def set_active_pool(pool: Optional[BufferPool]) -> Optional[BufferPool]:
    """route DeviceBuffer allocations / frees through `pool` (None: plain b2_dev_alloc / b2_dev_free); returns the
    previous setting"""
    global _ACTIVE_POOL
    prev, _ACTIVE_POOL = _ACTIVE_POOL, pool
    return prev


class DeviceBuffer:
    """cudaMalloc'd array of Fr elements (b2_dev_alloc), optionally recycled through the active BufferPool."""

    def __init__(self, elems: int):
        nbytes = max(1, elems) * 32
        pool = _ACTIVE_POOL
        got = pool.take(nbytes) if pool is not None else None
        if got is None:
            p = ctypes.c_void_p()
            rc = lib().b2_dev_alloc(nbytes, ctypes.byref(p))
            if rc and pool is not None and pool.cached_bytes:
                pool.trim()                     # out of memory with blocks parked in the pool: release them and retry
                rc = lib().b2_dev_alloc(nbytes, ctypes.byref(p))
            check(rc)
            got = p.value
        self.ptr, self.elems = got, elems
        self._pool, self._nbytes = pool, nbytes

    def upload(self, a: np.ndarray, offset_elems: int = 0) -> "DeviceBuffer":
        a = np.ascontiguousarray(a, dtype=np.uint64)
        check(lib().b2_memcpy_h2d(ctypes.c_void_p(self.ptr + offset_elems * 32), ctypes.c_void_p(a.ctypes.data), a.nbytes))
        return self

    def download(self, elems: Optional[int] = None, offset_elems: int = 0) -> np.ndarray:
        elems = self.elems - offset_elems if elems is None else elems
        out = np.empty((elems, 4), dtype=np.uint64)
        check(lib().b2_memcpy_d2h(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(self.ptr + offset_elems * 32), out.nbytes))
        return out

    def free(self):
        if self.ptr:
            if self._pool is not None:
                self._pool.give(self.ptr, self._nbytes)
            else:
                lib().b2_dev_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def coeff_to_extended_dev(domain, polys: np.ndarray, out: DeviceBuffer, out_col: int = 0) -> None:
    """EvaluationDomain::coeff_to_extended (poly/domain.rs:270-287) for a batch of host polynomials
    (columns, n, 4), written to device memory: out[(out_col + c) * 2^extended_k ...]."""
    polys = np.ascontiguousarray(polys, dtype=np.uint64)
    if polys.ndim == 2:
        polys = polys[None]
    z = np.concatenate([domain.g_coset, domain.g_coset_inv])
    d = NttDesc()
    d.log_n, d.location = domain.extended_k, 2
    d.omega = domain.extended_omega.ctypes.data
    d.coset_in = z.ctypes.data
    d.n_in = d.in_stride = domain.n
    d.n_out = d.out_stride = domain.extended_len()
    d.columns = polys.shape[0]
    d.in_ = polys.ctypes.data
    d.out = out.ptr + out_col * domain.extended_len() * 32
    check(lib().b2_ntt_exec(ctypes.byref(d)))


class Evaluator:
    """plonk::Evaluator (evaluation.rs:270-298) + the structure of the constraint system that
    evaluate_h reads (permutation columns and chunking, blinding factors, lookup / shuffle shapes)."""

    def __init__(self, rotations, constants, calculations, value_parts, lookup_results, shuffle_results,
                 num_fixed: int, num_advice: int, num_instance: int, permutation_columns, degree: int,
                 blinding_factors: int):
        self.rotations = list(rotations)
        self.constants = [c % _fr.R_MOD for c in constants]
        self.calculations = list(calculations)
        self.value_parts = list(value_parts)
        self.lookup_results = list(lookup_results)
        self.shuffle_results = list(shuffle_results)
        self.num_fixed, self.num_advice, self.num_instance = num_fixed, num_advice, num_instance
        self.permutation_columns = list(permutation_columns)
        self.chunk_len = degree - 2                       # evaluation.rs:1010
        self.blinding_factors = blinding_factors
        self._programs = {}

    # ---- program construction -----------------------------------------------------------
    def build_h_program(self, n_perm_sets: int, lookup_set_counts: Sequence[int], n_shuffles: int):
        f = self.flat_h_program(n_perm_sets, lookup_set_counts, n_shuffles)
        return QuotientProgram(f["rotations"], f["constants"], f["calcs"], f["result"], self.num_fixed, self.num_advice,
                               self.num_instance, f["n_aux"], f["n_challenges"])

    def flat_h_program(self, n_perm_sets: int, lookup_set_counts: Sequence[int], n_shuffles: int) -> dict:
        """Flat Calculation list for the whole of evaluate_h.  Aux columns, in order:
        l0, l_last, l_active_row, sigma cosets (one per permutation column), permutation z cosets (one per
        set), then per lookup its z cosets and its m coset, then the shuffle product cosets.
        Challenges: beta, gamma, theta, y, then beta * zeta * DELTA^j for every permutation column j."""
        rotations = list(self.rotations)
        constants = list(self.constants)
        calcs: List[tuple] = list(self.calculations)

        def rot_index(r: int) -> int:
            if r not in rotations:
                rotations.append(r)
            return rotations.index(r)

        def const(v: int):
            v %= _fr.R_MOD
            if v not in constants:
                constants.append(v)
            return ("Constant", constants.index(v))

        def emit(c: tuple):
            calcs.append(c)
            return ("Intermediate", len(calcs) - 1)

        rot_cur, rot_next = rot_index(0), rot_index(1)
        rot_last = rot_index(-(self.blinding_factors + 1))          # evaluation.rs:1009
        one = const(1)
        n_sigma = len(self.permutation_columns)
        AUX_L0, AUX_LLAST, AUX_LACTIVE = 0, 1, 2
        aux_sigma = 3
        aux_perm_z = aux_sigma + n_sigma
        aux_next = aux_perm_z + n_perm_sets

        def aux(i: int, rot: int = rot_cur):
            return ("Aux", i, rot)

        l0, l_last, l_active = aux(AUX_L0), aux(AUX_LLAST), aux(AUX_LACTIVE)
        acc = [None]

        def fold(term):                                             # *value = *value * y + term
            acc[0] = term if acc[0] is None else emit(("MulChAdd", acc[0], term, CH_Y))

        for part in self.value_parts:                               # :896-907
            fold(part)

        # permutation terms :1005-1090
        if n_perm_sets:
            zc = [aux(aux_perm_z + s) for s in range(n_perm_sets)]
            z_next = [aux(aux_perm_z + s, rot_next) for s in range(n_perm_sets)]
            z_lastrot = [aux(aux_perm_z + s, rot_last) for s in range(n_perm_sets)]
            fold(emit(("Mul", emit(("Sub", one, zc[0])), l0)))                                   # :1027-1028
            zl = zc[-1]
            fold(emit(("Mul", emit(("Sub", emit(("Mul", zl, zl)), zl)), l_last)))                # :1031-1035
            for s in range(1, n_perm_sets):                                                      # :1038-1046
                fold(emit(("Mul", emit(("Sub", zc[s], z_lastrot[s - 1])), l0)))
            for s in range(n_perm_sets):                                                         # :1053-1084
                cols = list(enumerate(self.permutation_columns))[s * self.chunk_len:(s + 1) * self.chunk_len]
                left = z_next[s]
                for j, (kind, idx) in cols:
                    v = (kind, idx, rot_cur)
                    t = emit(("Mul", ("Challenge", CH_BETA), aux(aux_sigma + j)))
                    t = emit(("AddChallenge", emit(("Add", v, t)), "Gamma"))
                    left = emit(("Mul", left, t))
                right = zc[s]
                for j, (kind, idx) in cols:
                    v = (kind, idx, rot_cur)
                    t = emit(("Mul", ("CosetX",), ("Challenge", CH_FIRST_DELTA + j)))
                    t = emit(("AddChallenge", emit(("Add", v, t)), "Gamma"))
                    right = emit(("Mul", right, t))
                fold(emit(("Mul", emit(("Sub", left, right)), l_active)))

        # lookups :1104-1180
        if len(lookup_set_counts) != len(self.lookup_results):
            raise B2Error(B2_ERR_ARG, "one set count per lookup expected")
        for res, sets_len in zip(self.lookup_results, lookup_set_counts):
            if sets_len != len(res[1]) or sets_len != len(res[2]):
                raise B2Error(B2_ERR_ARG, "lookup z-set count differs from the evaluator's input sets")
            z = [aux(aux_next + i) for i in range(sets_len)]
            z_next = [aux(aux_next + i, rot_next) for i in range(sets_len)]
            z_lastrot = [aux(aux_next + i, rot_last) for i in range(sets_len)]
            m = aux(aux_next + sets_len)
            aux_next += sets_len + 1
            table = emit(res[0])
            prod = [emit(c) for c in res[1]]
            sums = [emit(c) for c in res[2]]
            fold(emit(("Mul", z[0], l0)))                                                        # :1140
            fold(emit(("Mul", z[sets_len - 1], l_last)))                                         # :1143
            dz = emit(("Sub", z_next[0], z[0]))
            t = emit(("Mul", emit(("Add", emit(("Mul", dz, table)), m)), prod[0]))
            t = emit(("Sub", t, emit(("Mul", table, sums[0]))))
            fold(emit(("Mul", t, l_active)))                                                     # :1151-1156
            for i in range(1, sets_len):                                                         # :1159-1162
                fold(emit(("Mul", emit(("Sub", z[i], z_lastrot[i - 1])), l0)))
            for i in range(1, sets_len):                                                         # :1170-1176
                dz = emit(("Sub", z_next[i], z[i]))
                t = emit(("Sub", emit(("Mul", dz, prod[i])), sums[i]))
                fold(emit(("Mul", t, l_active)))

        # shuffles :1184-1220
        if n_shuffles != len(self.shuffle_results):
            raise B2Error(B2_ERR_ARG, "one product polynomial per shuffle group expected")
        for res in self.shuffle_results:
            z, z_next = aux(aux_next), aux(aux_next, rot_next)
            aux_next += 1
            inp, tab = emit(res[0]), emit(res[1])
            fold(emit(("Mul", emit(("Sub", one, z)), l0)))
            fold(emit(("Mul", emit(("Sub", emit(("Mul", z, z)), z)), l_last)))
            t = emit(("Sub", emit(("Mul", z_next, tab)), emit(("Mul", z, inp))))
            fold(emit(("Mul", t, l_active)))

        result = acc[0] if acc[0] is not None else const(0)
        return {"rotations": rotations, "constants": constants, "calcs": calcs, "result": result, "n_aux": aux_next,
                "n_challenges": CH_FIRST_DELTA + n_sigma}

    def program(self, n_perm_sets: int, lookup_set_counts: Sequence[int], n_shuffles: int) -> QuotientProgram:
        key = (n_perm_sets, tuple(lookup_set_counts), n_shuffles)
        if key not in self._programs:
            self._programs[key] = self.build_h_program(n_perm_sets, lookup_set_counts, n_shuffles)
        return self._programs[key]

    # ---- evaluate_h -----------------------------------------------------------------------
    def evaluate_h(self, domain, fixed_polys, advice_polys, instance_polys, l0, l_last, l_active_row, sigma_polys,
                   y: int, beta: int, gamma: int, theta: int, lookups, shuffles, permutations,
                   to_coeff: bool = False, zeta: Optional[int] = None, mode: str = "cosets",
                   cosets: Optional[Sequence[int]] = None) -> np.ndarray:
        """evaluation.rs:778-1226 for one proof.

        *_polys, sigma_polys, permutations (z per set), lookups ([{"z": [...], "m": poly}]), shuffles ([poly]):
        coefficient form, (n, 4) Montgomery arrays -- what the `cuda` variant of the reference takes
        (:1229-1241); l0 / l_last / l_active_row: extended-domain evaluations as ProvingKey holds them.
        Everything is extended on the device and stays there.  Returns the extended evaluations, or with
        to_coeff=True the coefficients of h(X) = numerator / (X^n - 1) (divide_by_vanishing_poly folded into
        the kernel's store, then extended_to_coeff on the device): vanishing/prover.rs:64-96.

        mode "extended": every polynomial is coset-extended to 2^extended_k rows first (the reference's
        layout).  mode "cosets" (default): the extended domain is walked as its 2^(extended_k - k)
        interleaved cosets zeta * extended_omega^c * <omega> -- extended row 2^(extended_k-k) * i + c is row i
        of coset c, and a rotation never leaves its coset -- so each pass needs only size-2^k transforms
        (no zero padding) and 2^-(extended_k-k) of the memory; `cosets` restricts the pass to some of them
        (the multi-GPU split: one coset per rank, rows of the others left untouched)."""
        require_gpu()
        ext_len = domain.extended_len()
        n = domain.n
        n_sets = len(permutations)
        set_counts = [len(lk["z"]) for lk in lookups]
        prog = self.program(n_sets, set_counts, len(shuffles))
        aux_polys = list(sigma_polys) + list(permutations)
        for lk in lookups:
            aux_polys += list(lk["z"]) + [lk["m"]]
        aux_polys += list(shuffles)
        groups = [list(fixed_polys), list(advice_polys), list(instance_polys), aux_polys]
        zeta_v = domain._zeta if zeta is None else zeta
        R = _fr.R_MOD
        challenges = [beta % R, gamma % R, theta % R, y % R]
        d = beta * zeta_v % R                                     # delta_start, evaluation.rs:1011
        for _ in self.permutation_columns:
            challenges.append(d)
            d = d * DELTA % R
        lag_host = [np.asarray(v, dtype=np.uint64).reshape(ext_len, 4) for v in (l0, l_last, l_active_row)]
        n_polys = sum(len(g) for g in groups)
        if mode == "extended":
            if cosets is not None:
                raise B2Error(B2_ERR_ARG, "`cosets` needs mode='cosets'")
            buf = DeviceBuffer((n_polys + 3) * ext_len)
            out = DeviceBuffer(ext_len)
            try:
                ptrs, col = [], 0
                for g in groups:
                    if g:
                        coeff_to_extended_dev(domain, np.stack([np.asarray(p, dtype=np.uint64).reshape(-1, 4) for p in g]),
                                              buf, col)
                    ptrs.append([buf.ptr + (col + i) * ext_len * 32 for i in range(len(g))])
                    col += len(g)
                lag = []
                for v in lag_host:
                    buf.upload(v, col * ext_len)
                    lag.append(buf.ptr + col * ext_len * 32)
                    col += 1
                scale = domain.t_evaluations if to_coeff else None
                prog.eval(domain.extended_k, 1 << (domain.extended_k - domain.k), ptrs[0], ptrs[1], ptrs[2],
                          lag + ptrs[3], challenges, out.ptr, x0=1, x_step=domain._ext_omega, scale=scale)
                return extended_to_coeff_dev(domain, out) if to_coeff else out.download()
            finally:
                buf.free()
                out.free()
        if mode != "cosets":
            raise B2Error(B2_ERR_ARG, f"unknown mode {mode!r}")
        n_cosets = 1 << (domain.extended_k - domain.k)
        which = list(range(n_cosets)) if cosets is None else list(cosets)
        if to_coeff and len(which) != n_cosets:
            raise B2Error(B2_ERR_ARG, "to_coeff needs every coset")
        out = DeviceBuffer(ext_len)
        try:
            self.evaluate_h_tasks(domain, groups, lag_host, challenges, prog, [(c, 0, n) for c in which], out.ptr,
                                  compact=False, scaled=to_coeff, zeta=zeta_v)
            return extended_to_coeff_dev(domain, out) if to_coeff else out.download()
        finally:
            out.free()

    def evaluate_h_tasks(self, domain, groups, lag_host, challenges, prog, tasks, out_ptr: int, compact: bool,
                         scaled: bool, zeta: int, resident: "Optional[ResidentPolys]" = None) -> None:
        """Coset-mode core.  tasks: (coset, row_begin, row_count) triples, coset-major; the rows of a task go to
        out[4 * row + coset] (compact=False, the extended layout) or are appended to `out` in task order
        (compact=True: what a rank contributes to the all-gather, parallel.sharded_evaluate_h).  scaled: fold
        divide_by_vanishing_poly into the store.  resident: polynomials already in HBM (ResidentPolys);
        otherwise they are uploaded for this call."""
        n = domain.n
        R = _fr.R_MOD
        n_cosets = 1 << (domain.extended_k - domain.k)
        res = resident if resident is not None else ResidentPolys(domain, groups, lag_host)
        try:
            loaded, written = None, 0
            for c, begin, count in tasks:
                if c != loaded:
                    g_c = zeta * pow(domain._ext_omega, c, R) % R
                    if res.n_polys:
                        coeff_to_coset_dev(domain, res.coef.ptr, res.n_polys, g_c, res.cos.ptr)
                    loaded = c
                scale = domain.t_evaluations[c:c + 1] if scaled else None
                if compact:
                    stride, offset = 1, written
                else:
                    stride, offset = n_cosets, c + begin * n_cosets
                p = res.ptrs
                prog.eval(domain.k, 1, p[0], p[1], p[2], res.lag_ptrs(c) + p[3], challenges, out_ptr,
                          x0=pow(domain._ext_omega, c, R), x_step=domain._omega, scale=scale,
                          out_stride=stride, out_offset=offset, row_begin=begin, row_count=count)
                written += count
        finally:
            if resident is None:
                res.free()


class ResidentPolys:
    """Everything evaluate_h reads, resident in HBM: the coefficient forms of all polynomials (one buffer,
    fixed | advice | instance | aux order), one coset's worth of scratch for their evaluations, and
    l0 / l_last / l_active_row split into the 2^(extended_k - k) cosets (they are proving-key data)."""

    def __init__(self, domain, groups, lag_host):
        n = domain.n
        nc = 1 << (domain.extended_k - domain.k)
        self.n_polys = sum(len(g) for g in groups)
        self.coef = DeviceBuffer(max(1, self.n_polys) * n)
        self.cos = DeviceBuffer(max(1, self.n_polys) * n)
        self.lag = DeviceBuffer(3 * nc * n)
        col = 0
        for g in groups:                       # column by column: no host-side concatenation of the whole set
            for p in g:
                self.coef.upload(np.asarray(p, dtype=np.uint64).reshape(n, 4), col * n)
                col += 1
        for c in range(nc):
            for i, v in enumerate(lag_host):
                v = np.asarray(v, dtype=np.uint64).reshape(-1, 4)
                self.lag.upload(np.ascontiguousarray(v[c::nc]), (c * 3 + i) * n)
        self.ptrs, col = [], 0
        for g in groups:
            self.ptrs.append([self.cos.ptr + (col + i) * n * 32 for i in range(len(g))])
            col += len(g)
        self._n = n

    def lag_ptrs(self, c: int):
        return [self.lag.ptr + (c * 3 + i) * self._n * 32 for i in range(3)]

    def free(self):
        self.coef.free()
        self.cos.free()
        self.lag.free()


def interleave_cosets_dev(domain, d_compact: int, d_ext: int) -> None:
    """coset-major h (coset c at d_compact[c * n ...]) -> extended layout d_ext[4 * i + c]: a zero-instruction
    program whose result is its one aux column, stored with the coset stride"""
    n_cosets = 1 << (domain.extended_k - domain.k)
    prog = QuotientProgram([0], [0], [], ("Aux", 0, 0), 0, 0, 0, 1, 0)
    try:
        for c in range(n_cosets):
            prog.eval(domain.k, 1, [], [], [], [d_compact + c * domain.n * 32], [], d_ext, out_stride=n_cosets,
                      out_offset=c)
    finally:
        prog.free()


def coeff_to_coset_dev(domain, d_coeffs: int, columns: int, gen: int, d_out: int, stream: int = 0) -> None:
    """Evaluations of `columns` device-resident polynomials (2^k coefficients each) on the coset gen * <omega>:
    one size-2^k transform per column with x[i] *= gen^i fused into its first pass (b2_ntt_desc.coset_gen).
    stream: a CUDA stream (b2_stream_create) makes the call asynchronous on it."""
    g = _fr.to_mont(gen)
    d = NttDesc()
    d.log_n, d.location = domain.k, 1
    d.omega = domain.omega.ctypes.data
    d.coset_gen = g.ctypes.data
    d.n_in = d.in_stride = d.n_out = d.out_stride = domain.n
    d.columns = columns
    d.in_, d.out = d_coeffs, d_out
    d.stream = stream or None
    check(lib().b2_ntt_exec(ctypes.byref(d)))


def extended_to_coeff_dev(domain, ext: DeviceBuffer) -> np.ndarray:
    """EvaluationDomain::extended_to_coeff (poly/domain.rs:328-350), device input, host output."""
    n_out = domain.n * domain.quotient_poly_degree
    out = np.empty((n_out, 4), dtype=np.uint64)
    z = np.concatenate([domain.g_coset_inv, domain.g_coset])   # leaving the coset: {zeta^2, zeta}
    d = NttDesc()
    d.log_n, d.location = domain.extended_k, 3
    d.omega = domain.extended_omega_inv.ctypes.data
    d.divisor = domain.extended_ifft_divisor.ctypes.data
    d.coset_out = z.ctypes.data
    d.n_in = d.in_stride = domain.extended_len()
    d.n_out = d.out_stride = n_out
    d.columns = 1
    d.in_ = ext.ptr
    d.out = out.ctypes.data
    check(lib().b2_ntt_exec(ctypes.byref(d)))
    return out
