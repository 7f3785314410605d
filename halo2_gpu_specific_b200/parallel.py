"""Multi-GPU MSM: one process per GPU, point ranges sharded across ranks.

Mirrors gpu_multiexp_bound (halo2_proofs/src/arithmetic.rs:413-440): part_len = ceil(n / N_GPU),
rank g owns points [g * part_len, (g + 1) * part_len) of the SRS (uploaded once, resident) and
the matching slice of every scalar vector; each rank returns one partial G1 point and the
partials are summed.  The reference sums them on the host inside one process; here the only
exchange is an all-gather of 96 bytes per rank over torch.distributed (NCCL on GPUs, gloo in
the CPU tests).  NTT / per-column commit batches shard by column with no collective at all
(`column_range`).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world() -> Tuple[int, int]:
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """arithmetic.rs:426: part_len = (n + n_gpu - 1) / n_gpu; chunks(part_len)"""
    part_len = (n + world_size - 1) // world_size
    lo = min(rank * part_len, n)
    return lo, min(lo + part_len, n)


def column_range(columns: int, world_size: int, rank: int) -> Tuple[int, int]:
    """independent per-column NTT / commit batches: contiguous blocks of columns per rank"""
    return shard_range(columns, world_size, rank)


def all_gather_partials(partial: np.ndarray) -> np.ndarray:
    """(12,) uint64 Jacobian partial of this rank -> (world, 12), rank order"""
    d = _dist()
    p = np.ascontiguousarray(np.asarray(partial, dtype=np.uint64).reshape(12))
    if d is None:
        return p.reshape(1, 12)
    import torch
    t = torch.from_numpy(p.view(np.int64).copy())
    backend = d.get_backend()
    if backend == "nccl":
        t = t.cuda()
    out = torch.empty(d.get_world_size() * 12, dtype=torch.int64, device=t.device)
    d.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(np.uint64).reshape(-1, 12)


def sharded_msm(scalars_shard: np.ndarray, srs_shard, max_bits: int = 254, *,
                local_msm: Optional[Callable] = None, combine: Optional[Callable] = None) -> np.ndarray:
    """MSM whose points are range-sharded over the ranks of the default process group.

    scalars_shard / srs_shard: this rank's slice (see shard_range).  Every rank returns the full
    result.  local_msm / combine default to the CUDA engine; the CPU tests inject the oracle to
    exercise the partition + gather logic without a GPU."""
    if local_msm is None or combine is None:
        from . import arithmetic
        local_msm = local_msm or arithmetic.gpu_multiexp_single_gpu_with_bound
        combine = combine or arithmetic.g1_sum
    partial = local_msm(scalars_shard, srs_shard, max_bits)
    _, ws = world()
    if ws == 1:
        return np.asarray(partial, dtype=np.uint64).reshape(12)
    return combine(all_gather_partials(partial))
