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


class ShardedMsmFuture:
    """sharded_msm in flight: the local partial is being computed; result() gathers and combines"""

    def __init__(self, fut, combine):
        self._fut, self._combine = fut, combine

    def result(self) -> np.ndarray:
        partial = self._fut.result()
        _, ws = world()
        if ws == 1:
            return np.asarray(partial, dtype=np.uint64).reshape(12)
        return self._combine(all_gather_partials(partial))


def sharded_msm_async(scalars_shard: np.ndarray, srs_shard, max_bits: int = 254) -> ShardedMsmFuture:
    """sharded_msm whose local part does not wait (arithmetic.gpu_multiexp_async): a caller that commits column after
    column keeps the upload of the next column under the kernels of this one.  Collectives happen in result(), so
    every rank must call result() on its futures in the same order."""
    from . import arithmetic
    return ShardedMsmFuture(arithmetic.gpu_multiexp_async(scalars_shard, srs_shard, max_bits), arithmetic.g1_sum)


# ---- evaluate_h over several GPUs -------------------------------------------------------------------
def quotient_tasks(n_cosets: int, n: int, world_size: int, rank: int):
    """The extended domain in coset-major order (coset 0 rows 0..n-1, coset 1, ...) is cut into world_size equal
    contiguous ranges; returns this rank's range as (coset, row_begin, row_count) triples.  With world_size <=
    n_cosets a rank owns whole cosets; beyond that the ranks sharing a coset each transform it and evaluate
    their rows.  Cosets never exchange data (a rotation stays inside its coset)."""
    total = n_cosets * n
    lo, hi = shard_range(total, world_size, rank)
    tasks = []
    while lo < hi:
        c, begin = divmod(lo, n)
        count = min(n - begin, hi - lo)
        tasks.append((c, begin, count))
        lo += count
    return tasks


def coset_transform_shares(n_cosets: int, world_size: int, rank: int, n_polys: int):
    """Ranks that evaluate rows of the same coset (world_size > n_cosets, see quotient_tasks) need the coset
    evaluations of ALL n_polys polynomials, but need not all compute them: the group splits the polynomials and swaps
    the shares point-to-point.  Returns (group_first_rank, [(rank, poly_lo, poly_hi), ...]) for this rank's group; a
    group of one (world_size <= n_cosets) transforms everything itself."""
    share = max(1, world_size // n_cosets)
    first = (rank // share) * share
    return first, [(first + j,) + shard_range(n_polys, share, j) for j in range(share)]


def all_gather_rows(local: "np.ndarray | object", rows_total: int):
    """every rank's compact slice of h (equal row counts) -> the full coset-major array, on every rank"""
    d = _dist()
    if d is None:
        return local
    import torch
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
    out = torch.empty((rows_total, 4), dtype=torch.int64, device=t.device)
    d.all_gather_into_tensor(out.view(-1), t.reshape(-1))
    return out if isinstance(local, torch.Tensor) else out.numpy().view(np.uint64)


def sharded_evaluate_h(ev, domain, fixed_polys, advice_polys, instance_polys, l0, l_last, l_active_row, sigma_polys,
                       y: int, beta: int, gamma: int, theta: int, lookups, shuffles, permutations, zeta=None,
                       resident=None) -> np.ndarray:
    """Evaluator::evaluate_h + h(X) with the rows of the extended domain split over the ranks of the default
    process group (one process per GPU): every rank holds the coefficient forms, evaluates its rows
    (quotient_tasks), the slices of h are all-gathered over NCCL (32 B per row, the only exchange), and each
    rank finishes with divide_by_vanishing_poly (folded into the kernel) + extended_to_coeff.
    Returns the coefficients of h(X), identical on every rank.  resident: an evaluation.ResidentPolys built from
    the same polynomials (then the *_polys arguments only give the counts): nothing is uploaded in the call."""
    import torch
    from . import _fr
    from .evaluation import DELTA, DeviceBuffer, extended_to_coeff_dev, interleave_cosets_dev
    rank, ws = world()
    n, ext_len = domain.n, domain.extended_len()
    nc = ext_len // n
    if (nc * n) % ws:
        raise ValueError("world size must divide the extended domain")
    prog = ev.program(len(permutations), [len(lk["z"]) for lk in lookups], len(shuffles))
    aux_polys = list(sigma_polys) + list(permutations)
    for lk in lookups:
        aux_polys += list(lk["z"]) + [lk["m"]]
    aux_polys += list(shuffles)
    groups = [list(fixed_polys), list(advice_polys), list(instance_polys), aux_polys]
    R = _fr.R_MOD
    zeta_v = domain._zeta if zeta is None else zeta
    challenges = [beta % R, gamma % R, theta % R, y % R]
    dlt = beta * zeta_v % R
    for _ in ev.permutation_columns:
        challenges.append(dlt)
        dlt = dlt * DELTA % R
    lag_host = [np.asarray(v, dtype=np.uint64).reshape(ext_len, 4) for v in (l0, l_last, l_active_row)]
    tasks = quotient_tasks(nc, n, ws, rank)
    mine = sum(t[2] for t in tasks)
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.empty((mine, 4), dtype=torch.int64, device=dev)
    ev.evaluate_h_tasks(domain, groups, lag_host, challenges, prog, tasks, local.data_ptr(), compact=True, scaled=True,
                        zeta=zeta_v, resident=resident)
    full = all_gather_rows(local, nc * n)
    torch.cuda.synchronize()
    ext = DeviceBuffer(ext_len)
    try:
        interleave_cosets_dev(domain, full.data_ptr(), ext.ptr)
        return extended_to_coeff_dev(domain, ext)
    finally:
        ext.free()
