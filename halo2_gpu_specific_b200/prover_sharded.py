"""create_proof on several GPUs, first step: the commitments.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  Every rank runs the SAME
`plonk.create_proof` with the same witness, proving key and rng seed, so every rank builds the same transcript; what is
divided is the multi-scalar multiplication work, which is most of a proof's device time (79 % for the benches/plonk.rs
circuit, profiles/r1_ncu_summary.md; the 24 full-width z-column commitments alone are 0.25 s of the 1.3 s zkWasm-shaped
proof): of every block of columns that is committed, rank r commits the contiguous share `parallel.column_range` gives
it -- the column-parallel option of SURVEY 8(e) ("replicate the SRS and give whole columns to GPUs as the reference's
pool does, plonk/prover.rs:293-299: no collective at all") -- and the resulting points (one per column) are
all-gathered so that each rank can hash all of them.  The witness itself is uploaded by every rank over its own PCIe
link and stays resident there, because the z columns and evaluate_h need all of it on every rank.

Status: the sharding and gather logic is validated on CPU with gloo (tests/test_parallel_cpu.py: proof bytes of a
2- and 3-rank run equal the single-process oracle proof) and on two B200s over NCCL (tools/sharded_proof_check.py,
profiles/r1_sharded_prover_2gpu.json: every rank's bytes equal the single-GPU proof; that run gathered the points with
all_gather_object, the fixed-size tensor gather that replaced it afterwards is covered with gloo).  At the size that fitted the
remaining GPU budget (k = 16, an 12 ms proof) the pickled all-gathers cost more than the divided MSMs save; measuring
it at zkWasm scale is round-2 work, and so is running the coset split of evaluate_h inside the prover on GPUs
(`ShardedQuotient` below: its division logic is covered on CPU with gloo, its device methods have not run yet).
"""
from __future__ import annotations

from typing import List, Optional

from . import _fr
from . import parallel
from .plonk import Point, ResidentEngine


class ShardedCommits:
    """Mixin for a block engine (put it before the engine class): commits of a block are divided over the ranks of the
    default process group by contiguous column ranges; every method returns the points of ALL columns, in column order,
    on every rank."""

    def _share(self, count: int):
        rank, world = parallel.world()
        return parallel.column_range(count, world, rank)

    @staticmethod
    def _gather(local: List[Point], count: int) -> List[Point]:
        """points of this rank's column range -> points of all `count` columns, in column order, on every rank: one
        all-gather of fixed-size records (flag, x, y as 9 x 64 bits), the same pattern as parallel.all_gather_partials"""
        d = parallel._dist()
        if d is None:
            return list(local)
        import numpy as np
        import torch
        world = d.get_world_size()
        part_len = (count + world - 1) // world                     # parallel.shard_range's stride
        rec = np.zeros((max(1, part_len), 9), dtype=np.uint64)
        for i, p in enumerate(local):
            if p is not None:
                rec[i, 0] = 1
                for l in range(4):
                    rec[i, 1 + l] = (p[0] >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
                    rec[i, 5 + l] = (p[1] >> (64 * l)) & 0xFFFFFFFFFFFFFFFF
        t = torch.from_numpy(rec.view(np.int64))
        if d.get_backend() == "nccl":
            t = t.cuda()
        out = torch.empty((world * rec.shape[0], 9), dtype=torch.int64, device=t.device)
        d.all_gather_into_tensor(out, t)
        rows = out.cpu().numpy().view(np.uint64).reshape(world, rec.shape[0], 9)
        points: List[Point] = []
        for c in range(count):
            r = rows[c // part_len, c % part_len]
            if not r[0]:
                points.append(None)
            else:
                points.append((sum(int(r[1 + l]) << (64 * l) for l in range(4)),
                               sum(int(r[5 + l]) << (64 * l) for l in range(4))))
        return points

    _inside = False       # True while this rank works on its own share: nested protocol calls pass straight through

    def _local(self):
        import contextlib

        @contextlib.contextmanager
        def scope():
            prev, self._inside = self._inside, True
            try:
                yield
            finally:
                self._inside = prev
        return scope()

    def commit_lagrange(self, block, max_bits: int = _fr.NUM_BITS) -> List[Point]:
        if self._inside:
            return super().commit_lagrange(block, max_bits)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = super().commit_lagrange(self.sub_block(block, lo, hi), max_bits) if hi > lo else []
        return self._gather(pts, self.block_count(block))

    def commit(self, block) -> List[Point]:
        if self._inside:
            return super().commit(block)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = super().commit(self.sub_block(block, lo, hi)) if hi > lo else []
        return self._gather(pts, self.block_count(block))

    def commit_lagrange_and_ifft(self, block) -> List[Point]:
        """own share: commitment + inverse transform in one pass; the other columns are needed in coefficient form on
        this rank too (evaluate_h reads every z polynomial), so they are only transformed"""
        if self._inside:
            return super().commit_lagrange_and_ifft(block)
        count = self.block_count(block)
        lo, hi = self._share(count)
        with self._local():
            pts = super().commit_lagrange_and_ifft(self.sub_block(block, lo, hi)) if hi > lo else []
            if lo > 0:
                self.lagrange_to_coeff(self.sub_block(block, 0, lo))
            if hi < count:
                self.lagrange_to_coeff(self.sub_block(block, hi, count))
        return self._gather(pts, count)

    def put_and_commit_lagrange(self, host, max_bits: Optional[int]):
        if self._inside:
            return super().put_and_commit_lagrange(host, max_bits)
        block = self.put(host)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = self.commit_columns_with_bound(self.sub_block(block, lo, hi), max_bits) if hi > lo else []
        return block, self._gather(pts, self.block_count(block))


class ShardedQuotient:
    """Mixin: evaluate_h divided over the ranks by cosets of the extended domain (a rotation never
    leaves its coset, so cosets exchange nothing).  Rank r evaluates cosets r, r + world, ... into a zeroed extended
    buffer; the buffers are summed across ranks (an integer all-reduce of disjoint supports: every row is written by
    exactly one rank) and every rank brings the sum back to the h(X) pieces.  The engine supplies
    evaluate_h_cosets / all_reduce_rows / h_pieces."""

    def evaluate_h_blocks(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta):
        rank, world = parallel.world()
        if world == 1:
            return super().evaluate_h_blocks(pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                                             y, beta, gamma, theta)
        dm = self.domain
        n_cosets = 1 << (dm.extended_k - dm.k)
        mine = [c for c in range(n_cosets) if c % world == rank]
        hext = self.evaluate_h_cosets(pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                                      y, beta, gamma, theta, mine)
        self.all_reduce_rows(hext)
        return self.h_pieces(hext)


class ShardedResidentEngine(ShardedCommits, ResidentEngine):
    """ResidentEngine whose commitments are shared out over the ranks (one process per GPU; call
    torch.cuda.set_device / _lib.set_device(local_rank) and init_process_group("nccl") first)"""


class ShardedResidentEngineQ(ShardedQuotient, ShardedCommits, ResidentEngine):
    """+ evaluate_h divided by cosets.  NOT YET RUN ON GPUS: the three device methods below restate
    ResidentEngine.evaluate_h_blocks with a coset subset, a zeroed output and an NCCL all-reduce; the division logic
    itself is covered on CPU (tests/test_parallel_cpu.py) through the test double."""

    def evaluate_h_cosets(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta, cosets):
        from .evaluation import coeff_to_coset_dev
        from .plonk import DELTA, DevBlock, R
        dm = self.domain
        n = dm.n
        nc = 1 << (dm.extended_k - dm.k)
        key_cosets = self._key_cosets(pk)
        F, S = pk.fixed_polys.shape[0], pk.sigma_polys.shape[0]
        prog = pk.ev.program(n_perm, list(lookup_z_counts), n_shuffles)
        challenges = [beta % R, gamma % R, theta % R, y % R]
        d = beta * dm._zeta % R
        for _ in range(S):
            challenges.append(d)
            d = d * DELTA % R
        witness = [b for b in (advice, instance, z_block, m_block) if b.count]
        cos = {id(b): self.alloc(b.count) for b in witness}
        hext = DevBlock(self._buffer(dm.extended_len()).ptr, 1, dm.extended_len())
        self._fr_vec(2, hext.ptr, hext.ptr, dm.extended_len(), hext.ptr)          # x - x = 0: a zeroed buffer
        ptrs = lambda b: [cos[id(b)].ptr + i * n * 32 for i in range(b.count)] if b.count else []     # noqa: E731
        for c in cosets:
            g_c = dm._zeta * pow(dm._ext_omega, c, R) % R
            for b in witness:
                coeff_to_coset_dev(dm, b.ptr, b.count, g_c, cos[id(b)].ptr)
            kc = key_cosets[c]
            kp = [kc.ptr + i * n * 32 for i in range(kc.count)]
            zp, mp = ptrs(z_block), ptrs(m_block)
            aux = kp[F + S:F + S + 3] + kp[F:F + S] + zp[:n_perm]
            pos = n_perm
            for li, cnt in enumerate(lookup_z_counts):
                aux += zp[pos:pos + cnt] + [mp[li]]
                pos += cnt
            aux += zp[pos:pos + n_shuffles]
            prog.eval(dm.k, 1, kp[:F], ptrs(advice), ptrs(instance), aux, challenges, hext.ptr,
                      x0=pow(dm._ext_omega, c, R), x_step=dm._omega, scale=dm.t_evaluations[c:c + 1],
                      out_stride=nc, out_offset=c)
        return hext

    def all_reduce_rows(self, hext) -> None:
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        from .plonk import _DevArray
        check(lib().b2_synchronize())
        t = torch.as_tensor(_DevArray(hext.ptr, (hext.n, 4)), device=torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()

    def h_pieces(self, hext):
        import ctypes
        import numpy as np
        from ._lib import NttDesc, check, lib
        dm = self.domain
        pieces = dm.quotient_poly_degree
        hcoef = self.alloc(pieces)
        z = np.concatenate([dm.g_coset_inv, dm.g_coset])
        t = NttDesc()
        t.log_n, t.location = dm.extended_k, 1
        t.omega, t.divisor = dm.extended_omega_inv.ctypes.data, dm.extended_ifft_divisor.ctypes.data
        t.coset_out = z.ctypes.data
        t.n_in = t.in_stride = dm.extended_len()
        t.n_out = t.out_stride = dm.n * pieces
        t.columns, t.in_, t.out = 1, hext.ptr, hcoef.ptr
        check(lib().b2_ntt_exec(ctypes.byref(t)))
        return hcoef
