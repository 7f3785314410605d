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
profiles/r1_sharded_prover_2gpu.json: every rank's bytes equal the single-GPU proof).  At the size that fitted the
remaining GPU budget (k = 16, an 12 ms proof) the pickled all-gathers cost more than the divided MSMs save; measuring
it at zkWasm scale, and splitting evaluate_h by cosets inside the prover (parallel.sharded_evaluate_h has that split
for host inputs), are round-2 work.
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
    def _gather(local: List[Point]) -> List[Point]:
        d = parallel._dist()
        if d is None:
            return list(local)
        parts: list = [None] * d.get_world_size()
        d.all_gather_object(parts, list(local))
        return [p for part in parts for p in part]

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
        return self._gather(pts)

    def commit(self, block) -> List[Point]:
        if self._inside:
            return super().commit(block)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = super().commit(self.sub_block(block, lo, hi)) if hi > lo else []
        return self._gather(pts)

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
        return self._gather(pts)

    def put_and_commit_lagrange(self, host, max_bits: Optional[int]):
        if self._inside:
            return super().put_and_commit_lagrange(host, max_bits)
        block = self.put(host)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = self.commit_columns_with_bound(self.sub_block(block, lo, hi), max_bits) if hi > lo else []
        return block, self._gather(pts)


class ShardedResidentEngine(ShardedCommits, ResidentEngine):
    """ResidentEngine whose commitments are shared out over the ranks (one process per GPU; call
    torch.cuda.set_device / _lib.set_device(local_rank) and init_process_group("nccl") first)"""
