"""Test double for halo2_gpu_specific_b200.plonk.Engine: the same method set answered by the CPU oracle
(oracle/plonk.py, oracle/bn254.py, oracle/cpu_ref.c).  It exists so that the HOST logic of the prover mirror --
transcript order, RNG order, query grouping, multiplicities, keygen bookkeeping -- can be checked on CPU against
oracle/prover.py without a GPU.  It lives in tests/ because only tests may touch the oracle; the engines of the
package itself (ResidentEngine, Engine, the sharded variants) all run on the device and refuse to start without one."""
from __future__ import annotations

import numpy as np

from oracle import bn254 as o
from oracle import cref
from oracle import plonk as P

from halo2_gpu_specific_b200.plonk import ArrayBlocks

R = o.R_MOD
enc, dec = o.fr_encode, o.fr_decode


class _Blinds:
    """hands pre-drawn blinding values to oracle/plonk.py's rng.randrange(R) calls, in order"""

    def __init__(self, blocks):
        self.q = [v for b in blocks for v in dec(b)]

    def randrange(self, m):
        assert m == R
        return self.q.pop(0)


class OracleEngine(ArrayBlocks):
    """array primitives answered by the oracle; the block protocol on top of them is the package's own ArrayBlocks
    (the same code the host-API device engine uses)"""

    def __init__(self, oracle_params, oracle_domain, oracle_cs):
        self.p, self.d, self.cs = oracle_params, oracle_domain, oracle_cs
        self.domain = oracle_domain
        self.ev = P.Evaluator.new(oracle_cs)

    # -- commitments
    def _msm(self, col, bases):
        col = np.ascontiguousarray(col, dtype=np.uint64).reshape(-1, 4)
        j = cref.best_multiexp(col, bases[:col.shape[0]])
        return o.g1_affine_decode(cref.jac_to_affine(j))[0]

    def commit_lagrange(self, cols, max_bits=254):
        for c in cols:
            assert all(v.bit_length() <= max_bits for v in dec(c)), "commit_lagrange_with_bound: scalar above the bound"
        return [self._msm(c, self.p.g_lagrange) for c in cols]

    def commit_lagrange_and_ifft(self, cols):
        pts = [self._msm(c, self.p.g_lagrange) for c in cols]
        self.lagrange_to_coeff(cols)
        return pts

    def commit(self, cols):
        return [self._msm(c, self.p.g) for c in cols]

    # -- transforms
    def lagrange_to_coeff(self, cols):
        d = self.d
        for i in range(cols.shape[0]):
            cols[i] = cref.ifft(cols[i], enc([d.omega_inv])[0], enc([d.ifft_divisor])[0], d.k)
        return cols

    def coeff_to_extended(self, cols):
        d = self.d
        z, z2, w = enc([d.g_coset])[0], enc([d.g_coset_inv])[0], enc([d.extended_omega])[0]
        return np.stack([cref.coeff_to_extended(c, d.k, d.extended_k, z, z2, w) for c in cols])

    def fft(self, a):
        a[:] = cref.best_fft(a, enc([self.d.omega])[0], self.d.k)
        return a

    # -- element-wise
    def fr_vec(self, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        return cref.field_vec(0, {"mul": 0, "add": 1, "sub": 2}[op], a, b)

    def to_mont(self, canonical):
        return cref.to_mont(0, np.ascontiguousarray(canonical, dtype=np.uint64).reshape(-1, 4))

    def from_mont(self, a):
        return cref.from_mont(0, np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4))

    # -- z columns and the quotient
    @staticmethod
    def _cols(advice, fixed, instance):
        return [dec(c) for c in advice], [dec(c) for c in fixed], [dec(c) for c in instance]

    def compress(self, expression_lists, advice, fixed, instance, theta):
        a, f, i = self._cols(advice, fixed, instance)
        n = self.d.n
        return np.stack([enc(P.evaluate_with_theta(ex, n, 1, f, a, i, theta)) for ex in expression_lists])

    def permutation_commit(self, cs, sigmas, advice, fixed, instance, beta, gamma, blinds):
        a, f, i = self._cols(advice, fixed, instance)
        zs = P.permutation_commit(self.cs, self.d, [dec(s) for s in sigmas], a, f, i, beta, gamma, _Blinds(blinds))
        return [enc(z) for z in zs]

    def logup_commit_z(self, cs, lookup, advice, fixed, instance, m, theta, beta):
        a, f, i = self._cols(advice, fixed, instance)
        n = self.d.n
        comp = lambda ex: P.evaluate_with_theta(ex, n, 1, f, a, i, theta)            # noqa: E731
        input_sets = [[comp(inp) for inp in s] for s in lookup["input_expressions_sets"]]
        table = comp(lookup["table_expressions"])
        return [enc(z) for z in P.logup_commit_z(self.cs, self.d, input_sets, table, dec(m), beta)]

    def shuffle_commit_product(self, cs, group, advice, fixed, instance, theta, beta):
        a, f, i = self._cols(advice, fixed, instance)
        return enc(P.shuffle_commit_product(self.cs, self.d, group, theta, beta, a, f, i))

    def evaluate_h(self, pk, advice_polys, instance_polys, y, beta, gamma, theta, lookups, shuffles, permutations):
        d = self.d
        ext = lambda p: d.coeff_to_extended(dec(p))                                 # noqa: E731
        h = P.evaluate_h(self.ev, self.cs, d, [ext(p) for p in pk.fixed_polys], [ext(p) for p in advice_polys],
                         [ext(p) for p in instance_polys], dec(pk.l0), dec(pk.l_last), dec(pk.l_active_row),
                         [ext(p) for p in pk.sigma_polys], y, beta, gamma, theta,
                         [{"z_cosets": [ext(z) for z in lk["z"]], "m_coset": ext(lk["m"])} for lk in lookups],
                         [ext(p) for p in shuffles], [ext(p) for p in permutations])
        return enc(d.extended_to_coeff(d.divide_by_vanishing_poly(h)))

    # -- evaluation and opening
    def eval_polynomial(self, poly, point):
        return o.eval_polynomial(dec(poly), point)

    def poly_combine(self, polys, v):
        acc = [0] * len(polys[0])
        for p in polys:
            acc = [(x * v + y) % R for x, y in zip(acc, dec(p))]
        return enc(acc)

    def kate_division(self, poly, z):
        return enc(o.kate_division(dec(poly), z))


class CosetQuotientDouble:
    """evaluate_h_blocks(..., tasks=, combine=) / all_reduce_rows for the oracle-backed engine (what
    prover_sharded.ShardedQuotient asks of an engine): the oracle evaluates the whole extended domain, rows outside
    this rank's (coset, row_begin, row_count) tasks are zeroed, the sum over ranks goes through torch.distributed on
    the host (gloo)"""

    def evaluate_h_blocks(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta, tasks=None, combine=None):
        if tasks is None:
            return super().evaluate_h_blocks(pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts,
                                             n_shuffles, y, beta, gamma, theta)
        d = self.d
        ext = lambda p: d.coeff_to_extended(dec(p))                                 # noqa: E731
        lookups, pos = [], n_perm
        for li, cnt in enumerate(lookup_z_counts):
            lookups.append({"z_cosets": [ext(z_block[pos + i]) for i in range(cnt)], "m_coset": ext(m_block[li])})
            pos += cnt
        h = P.evaluate_h(self.ev, self.cs, d, [ext(p) for p in pk.fixed_polys], [ext(p) for p in advice],
                         [ext(p) for p in instance], dec(pk.l0), dec(pk.l_last), dec(pk.l_active_row),
                         [ext(p) for p in pk.sigma_polys], y, beta, gamma, theta, lookups,
                         [ext(z_block[pos + i]) for i in range(n_shuffles)], [ext(z_block[i]) for i in range(n_perm)])
        h = d.divide_by_vanishing_poly(h)
        nc = 1 << (d.extended_k - d.k)
        own = {(b + j) * nc + c for c, b, cnt in tasks for j in range(cnt)}          # row i of coset c sits at i * nc + c
        hext = enc([v if i in own else 0 for i, v in enumerate(h)])
        combine(hext)
        coeffs = d.extended_to_coeff(dec(hext))
        n = d.n
        pieces = len(coeffs) // n
        return np.ascontiguousarray(enc(coeffs[:pieces * n])).reshape(pieces, n, 4)

    def all_reduce_rows(self, hext):
        import torch.distributed as dist
        dist.all_reduce(self.as_tensor(hext), op=dist.ReduceOp.SUM)

    # -- what ShardedCommits.put_and_commit_lagrange / exchange_columns ask for (gloo on host arrays)
    @staticmethod
    def as_tensor(block):
        import torch
        return torch.from_numpy(block.view(np.int64))

    @staticmethod
    def before_collective():
        pass

    @staticmethod
    def after_collective():
        pass

    # -- what ShardedCommits.commit_by_point_range asks for: partial MSMs over a point range, and their sum
    def msm_partials(self, basis, block, lo, hi, max_bits):
        import torch
        bases = getattr(self.p, basis)
        out = np.zeros((block.shape[0], 12), dtype=np.uint64)
        for i, col in enumerate(block):
            out[i] = cref.best_multiexp(np.ascontiguousarray(col[lo:hi]), bases[lo:hi]) if hi > lo else \
                o.g1_jacobian_encode(None)
        self.partial_ranges = getattr(self, "partial_ranges", []) + [(lo, hi, block.shape[0])]
        return torch.from_numpy(out.view(np.int64))

    def sum_partials(self, gathered, world, count):
        rows = gathered.numpy().view(np.uint64).reshape(world, count, 12)
        return [o.g1_affine_decode(cref.jac_to_affine(cref.jac_sum(np.ascontiguousarray(rows[:, c]))))[0]
                for c in range(count)]

    def put_share_and_commit(self, block, host, lo, hi, max_bits):
        """only this rank's columns are copied in: the others stay zero until exchange_columns delivers them"""
        block[lo:hi] = host[lo:hi]
        return self.commit_columns_with_bound(block[lo:hi], max_bits)
