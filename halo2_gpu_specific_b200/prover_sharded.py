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

Blocks with fewer columns than the ranks can share evenly -- the instance polynomial, the vanishing argument's random
polynomial, the h(X) pieces, the multiopen witness polynomials: the "few big MSMs" of SURVEY 8(e) -- are divided the
other way, option (i) there and what gpu_multiexp_bound does inside one process (arithmetic.rs:413-440): every rank
multiplies the point range [rank * ceil(n / N), ...) of EVERY column of the block, the Jacobian partials (96 bytes per
column and rank) are all-gathered in one collective and summed (`commit_by_point_range`; the choice between the two
divisions is a cost model of the measured MSM times, `_by_point_range`).

Status: the sharding and gather logic of both divisions is validated on CPU with gloo (tests/test_parallel_cpu.py:
proof bytes of a 2- and 3-rank run equal the single-process oracle proof) and on 2, 4 and 8 B200s over NCCL
(tools/sharded_proof_check.py, profiles/r2_sharded_proof_*.json: every rank's bytes equal the single-GPU proof at
k = 18 .. 22, coset split of evaluate_h included); `bench.py` at N > 1 proves a k = 18 circuit with all ranks together
and compares every rank's bytes with the single-GPU proof (`sharded_create_proof`), which is how the driver sees it.
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

    EARLY_TRANSFORMS = False    # the ranks divide the advice columns and their transforms among themselves instead

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

    # -- few big MSMs: divided by point range (SURVEY 8(e) option (i); arithmetic.rs:413-440)
    RANGE_SHARD: Optional[bool] = None      # None: by the cost model below; True / False: always / never (tests, A/B runs)

    @staticmethod
    def _msm_ms(points: int) -> float:
        """one MSM of `points` full-width scalars on one B200, ms: fit of the measured sweep (2^18: 1.20, 2^20: 3.22,
        2^22: 10.57; profiles/r2_bench_1gpu_final.json strong_scaling) -- a fixed sort / bucket-reduction part plus the
        point additions"""
        return 0.6 + 10.0 * points / (1 << 22)

    def _by_point_range(self, count: int, max_bits: int) -> bool:
        """divide the `count` columns of a block by point range instead of by column?  Column-parallel costs
        ceil(count / N) whole MSMs on the busiest rank, the range division `count` MSMs of n / N points on every rank:
        it wins when ranks would idle (count < N: one instance column, the random polynomial, h pieces on 8 ranks) and
        loses when the fixed part of an MSM, paid `count` times, outweighs the idle ranks.  Bounded commits stay
        column-parallel (their bound check lives in the batch call)."""
        _, world = parallel.world()
        if world == 1 or count == 0 or max_bits < _fr.NUM_BITS or not hasattr(self, "msm_partials"):
            return False
        if self.RANGE_SHARD is not None:
            return bool(self.RANGE_SHARD)
        n = self.domain.n
        return count * self._msm_ms(-(-n // world)) < -(-count // world) * self._msm_ms(n)

    def commit_by_point_range(self, basis: str, block, max_bits: int = _fr.NUM_BITS) -> List[Point]:
        """every column of `block` against params.<basis> ("g" or "g_lagrange") with the POINTS divided over the ranks:
        rank r multiplies scalars and bases [r * ceil(n / N), (r + 1) * ceil(n / N)) of every column
        (parallel.shard_range, the reference's part_len rule), ONE all-gather moves the N x count Jacobian partials
        (rank-major, 96 bytes each) and every rank adds them up.  The engine supplies msm_partials (block, range ->
        (count, 12) int64 tensor of un-normalised Jacobian points) and sum_partials ((N * count, 12) rank-major ->
        points)."""
        d = parallel._dist()
        import torch
        rank, world = parallel.world()
        count, n = self.block_count(block), self.domain.n
        lo, hi = parallel.shard_range(n, world, rank)
        with self._local():
            part = self.msm_partials(basis, block, lo, hi, max_bits)
        if tuple(part.shape) != (count, 12) or part.dtype != torch.int64:
            raise ValueError("msm_partials must return a (count, 12) int64 tensor")
        gathered = torch.empty((world * count, 12), dtype=torch.int64, device=part.device)
        self.before_collective()
        d.all_gather_into_tensor(gathered, part.contiguous())
        self.after_collective()
        self.range_commits = getattr(self, "range_commits", 0) + count
        return self.sum_partials(gathered, world, count)

    def commit_lagrange(self, block, max_bits: int = _fr.NUM_BITS) -> List[Point]:
        if self._inside:
            return super().commit_lagrange(block, max_bits)
        if self._by_point_range(self.block_count(block), max_bits):
            return self.commit_by_point_range("g_lagrange", block, max_bits)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = super().commit_lagrange(self.sub_block(block, lo, hi), max_bits) if hi > lo else []
        return self._gather(pts, self.block_count(block))

    def commit(self, block) -> List[Point]:
        if self._inside:
            return super().commit(block)
        if self._by_point_range(self.block_count(block), _fr.NUM_BITS):
            return self.commit_by_point_range("g", block)
        lo, hi = self._share(self.block_count(block))
        with self._local():
            pts = super().commit(self.sub_block(block, lo, hi)) if hi > lo else []
        return self._gather(pts, self.block_count(block))

    def commit_lagrange_and_ifft(self, block) -> List[Point]:
        """own share: commitment + inverse transform in one pass; the other columns arrive from their owners in
        coefficient form (exchange_columns) -- evaluate_h reads every z polynomial on every rank.  The owners are also
        the only ranks that built those columns (owns_columns)."""
        if self._inside:
            return super().commit_lagrange_and_ifft(block)
        count = self.block_count(block)
        lo, hi = self._share(count)
        with self._local():
            pts = super().commit_lagrange_and_ifft(self.sub_block(block, lo, hi)) if hi > lo else []
        self.exchange_columns(block)
        return self._gather(pts, count)

    def owns_columns(self, block, lo: int, hi: int) -> bool:
        """does this rank commit (hence build) any of the columns [lo, hi) of `block`?"""
        mine_lo, mine_hi = self._share(self.block_count(block))
        return lo < mine_hi and mine_lo < hi

    def lagrange_to_coeff(self, block):
        """wide blocks (the advice columns): every rank transforms its share and the shares are exchanged over
        NVLink; narrow blocks are cheaper to transform everywhere than to move"""
        _, world = parallel.world()
        count = self.block_count(block)
        if self._inside or world == 1 or count < 4 * world:
            return super().lagrange_to_coeff(block)
        lo, hi = self._share(count)
        with self._local():
            if hi > lo:
                super().lagrange_to_coeff(self.sub_block(block, lo, hi))
        self.exchange_columns(block)
        return block

    def multiplicity_block(self, cs, pk, advice, instance, theta: int, blinds, only=None):
        """the lookups are divided like the columns of the m block (so a rank counts exactly the m columns it will
        commit), the columns exchanged and the commit bound maximised over the ranks"""
        d = parallel._dist()
        n_lookups = len(cs.lookups)
        if self._inside or d is None or n_lookups == 0:
            return super().multiplicity_block(cs, pk, advice, instance, theta, blinds, only)
        lo, hi = self._share(n_lookups)
        with self._local():
            ms, bits = super().multiplicity_block(cs, pk, advice, instance, theta, blinds, only=range(lo, hi))
        self.exchange_columns(ms)
        import torch
        t = torch.tensor([bits], dtype=torch.int64)
        if d.get_backend() == "nccl":
            t = t.cuda()
        d.all_reduce(t, op=d.ReduceOp.MAX)
        return ms, int(t.item())

    def put_and_commit_lagrange(self, host, max_bits: Optional[int]):
        """The advice columns: rank r brings ONLY its share of the columns across its PCIe link (committing them on
        the way in, copy of column i + 1 under the MSM of column i), then the ranks hand each other their shares over
        NVLink (`exchange_columns`: one broadcast per rank), because evaluate_h and the z columns read every advice
        column on every rank.  Host->device traffic per rank drops from the whole witness to 1/world of it.  Engines
        without put_share_and_commit (the host-API engine) upload everything and divide only the commitments."""
        if self._inside:
            return super().put_and_commit_lagrange(host, max_bits)
        count = host.shape[0]
        lo, hi = self._share(count)
        _, world = parallel.world()
        if world > 1 and hasattr(self, "put_share_and_commit"):
            block = self.alloc(count)
            with self._local():
                pts = self.put_share_and_commit(block, host, lo, hi, max_bits) if hi > lo else []
            self.exchange_columns(block)
            return block, self._gather(pts, count)
        block = self.put(host)
        with self._local():
            pts = self.commit_columns_with_bound(self.sub_block(block, lo, hi), max_bits) if hi > lo else []
        return block, self._gather(pts, count)

    def exchange_columns(self, block) -> None:
        """every rank's column range of `block` (parallel.column_range) -> every rank: one broadcast per owning rank"""
        d = parallel._dist()
        if d is None:
            return
        count, world = self.block_count(block), d.get_world_size()
        self.before_collective()
        for r in range(world):
            lo, hi = parallel.column_range(count, world, r)
            if hi > lo:
                d.broadcast(self.as_tensor(self.sub_block(block, lo, hi)), src=r)
        self.after_collective()


class ShardedQuotient:
    """Mixin: evaluate_h divided over the ranks by contiguous row ranges of the coset-major extended domain
    (parallel.quotient_tasks; SURVEY 8e, DESIGN 5: a rotation never leaves its coset, so cosets exchange nothing).
    Rank r evaluates its rows into a zeroed extended buffer; the buffers are summed across ranks (an integer
    all-reduce of disjoint supports: every row is written by exactly one rank, so no limb ever carries) and every
    rank brings the sum back to the h(X) pieces.  The engine supplies evaluate_h_blocks(..., tasks=, combine=)."""

    def early_coset_ids(self):
        """the cosets this rank evaluates rows of (what its early transforms of the advice share are good for)"""
        rank, world = parallel.world()
        dm = self.domain
        return [c for c, _, _ in parallel.quotient_tasks(1 << (dm.extended_k - dm.k), dm.n, world, rank)]

    def evaluate_h_blocks(self, pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                          y, beta, gamma, theta):
        rank, world = parallel.world()
        if world == 1:
            return super().evaluate_h_blocks(pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                                             y, beta, gamma, theta)
        dm = self.domain
        tasks = parallel.quotient_tasks(1 << (dm.extended_k - dm.k), dm.n, world, rank)
        return super().evaluate_h_blocks(pk, advice, instance, z_block, m_block, n_perm, lookup_z_counts, n_shuffles,
                                         y, beta, gamma, theta, tasks=tasks, combine=self.all_reduce_rows)


class _ResidentCollectives:
    """what the sharding mixins ask of the device-resident engine: device blocks as torch tensors (no copy), the
    engine's lanes drained before NCCL touches the memory and NCCL's stream drained after"""

    @staticmethod
    def as_tensor(block):
        import torch
        from .plonk import _DevArray
        return torch.as_tensor(_DevArray(block.ptr, (max(1, block.count) * block.n, 4)),
                               device=torch.device("cuda", torch.cuda.current_device()))

    @staticmethod
    def before_collective() -> None:
        from ._lib import check, lib
        check(lib().b2_synchronize())

    @staticmethod
    def after_collective() -> None:
        import torch
        torch.cuda.synchronize()

    def put_share_and_commit(self, block, host, lo: int, hi: int, max_bits: Optional[int]):
        """columns [lo, hi) of `host` -> the same columns of the resident `block`, committed on the way in"""
        import numpy as np
        from ._lib import B2_ERR_ARG, B2Error
        if not (host.flags.c_contiguous and host.dtype == np.uint64 and host.ndim == 3):
            raise B2Error(B2_ERR_ARG, "expected a C-contiguous uint64 (columns, n, 4) array")
        bits = 0xFFFFFFFF if max_bits is None else max_bits
        dm = self.domain
        nc = 1 << (dm.extended_k - dm.k)
        ids = sorted(set(self.early_coset_ids())) if hasattr(self, "early_coset_ids") else list(range(nc))
        if self.EARLY_SHARE_TRANSFORMS and hi - lo >= 4 and (len(ids) + 1) * block.count * dm.n * 32 <= self.EARLY_TRANSFORM_BYTES:
            # the rank's link is busy for as long as its share is on the way and its multiplier idles: the coefficient
            # forms of the share (what lagrange_to_coeff would make later) and the share's evaluations on the cosets this
            # rank will work on in evaluate_h are made behind the upload (ResidentEngine.put_columns_with_early_transforms)
            return self.put_columns_with_early_transforms(block, host, lo, hi, bits, ids)
        own = self.sub_block(block, lo, hi)
        return self._commit(self.params.g_lagrange, host[lo:hi].ctypes.data, own, bits, False)

    EARLY_SHARE_TRANSFORMS = True

    def msm_partials(self, basis: str, block, lo: int, hi: int, max_bits: int):
        """points [lo, hi) of every column of the resident `block` against the same range of params.<basis> (the window
        table covers the whole SRS, b2_msm_dev takes the offset): one asynchronous MSM per column on the library's
        lanes, results (Jacobian, not normalised) left in device memory for the all-gather"""
        import ctypes
        import torch
        from ._lib import check, lib
        srs = getattr(self.params, basis)
        vp, L = ctypes.c_void_p, lib()
        out = torch.empty((block.count, 12), dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
        for i in range(block.count):
            check(L.b2_msm_dev(srs.handle, lo, vp(block.ptr + (i * block.n + lo) * 32), hi - lo, int(max_bits),
                               vp(out.data_ptr() + 96 * i), None))
        return out

    def sum_partials(self, gathered, world: int, count: int) -> List[Point]:
        """(world * count, 12) rank-major partials on the device -> the `count` points: one launch
        (b2_g1_sum_groups_dev), 96 bytes per column read back and normalised on the host as every commit entry does"""
        import ctypes
        import numpy as np
        import torch
        from ._lib import check, lib
        from .plonk import _points
        vp, L = ctypes.c_void_p, lib()
        sums = torch.empty((count, 12), dtype=torch.int64, device=gathered.device)
        check(L.b2_g1_sum_groups_dev(vp(gathered.data_ptr()), world, count, vp(sums.data_ptr()), None))
        check(L.b2_synchronize())
        host = np.ascontiguousarray(sums.cpu().numpy().view(np.uint64).reshape(count, 12))
        check(L.b2_g1_normalize(vp(host.ctypes.data), count))
        return _points(host)

    def all_reduce_rows(self, hext) -> None:
        import torch.distributed as dist
        self.before_collective()
        dist.all_reduce(self.as_tensor(hext), op=dist.ReduceOp.SUM)
        self.after_collective()


class ShardedResidentEngine(ShardedCommits, _ResidentCollectives, ResidentEngine):
    """ResidentEngine whose commitments are shared out over the ranks (one process per GPU; call
    torch.cuda.set_device / _lib.set_device(local_rank) and init_process_group("nccl") first)"""


class ShardedResidentEngineQ(ShardedQuotient, ShardedCommits, _ResidentCollectives, ResidentEngine):
    """+ evaluate_h divided by rows of the extended domain: ResidentEngine.evaluate_h_blocks does the work for the
    rank's (coset, row range) tasks; the NCCL all-reduce (_ResidentCollectives.all_reduce_rows) completes the buffer."""


# ---- every rank must draw the same randomness ------------------------------------------------------------------
def synchronized_rng():
    """The rng of a multi-rank proof: rank 0 draws a 256-bit seed from the OS and broadcasts it; every rank expands
    it with BLAKE2b in counter mode (plonk.Blake2bRng), so all ranks blind advice / m / z / h identically -- the
    precondition of ShardedCommits, where a rank commits only its share of columns that every rank evaluates."""
    import os
    import numpy as np
    from .plonk import Blake2bRng
    d = parallel._dist()
    if d is None:
        return Blake2bRng()
    import torch
    t = torch.from_numpy(np.frombuffer(os.urandom(32), dtype=np.uint8).copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.broadcast(t, src=0)
    return Blake2bRng(t.cpu().numpy().tobytes())


def assert_ranks_agree(data: bytes, what: str = "proof") -> None:
    """all-gather a 32-byte BLAKE2b digest of `data`; raise on every rank when any two differ (a rank that blinded
    differently, or was given another witness, produces a transcript no verifier accepts)"""
    import hashlib
    import numpy as np
    from ._lib import B2_ERR_ARG, B2Error
    d = parallel._dist()
    if d is None:
        return
    import torch
    t = torch.from_numpy(np.frombuffer(hashlib.blake2b(data, digest_size=32).digest(), dtype=np.uint8).copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    out = torch.empty(d.get_world_size() * 32, dtype=torch.uint8, device=t.device)
    d.all_gather_into_tensor(out, t)
    rows = out.cpu().numpy().reshape(-1, 32)
    if not (rows == rows[0]).all():
        raise B2Error(B2_ERR_ARG, f"multi-GPU create_proof: the ranks produced different {what} bytes (every rank "
                                  "needs the same witness, proving key and rng stream: use synchronized_rng())")


def create_proof(params, pk, advice, instances, rng=None, *, engine=None, split_quotient: bool = True, **kw) -> bytes:
    """plonk.create_proof over the ranks of the default process group: commitments divided by columns, evaluate_h by
    rows of the extended domain.  rng: None = synchronized_rng(); a caller-supplied rng must produce the same
    stream on every rank.  The proof is returned only after the ranks' bytes were found equal."""
    from . import plonk
    own = engine is None
    if own:
        engine = (ShardedResidentEngineQ if split_quotient else ShardedResidentEngine)(params, pk.vk.domain)
    try:
        proof = plonk.create_proof(params, pk, advice, instances, rng if rng is not None else synchronized_rng(),
                                   engine=engine, **kw)
    finally:
        if own:
            engine.free()
    assert_ranks_agree(proof)
    return proof
