"""Witness file of halo2_proofs/src/helpers.rs:919-1015 (AssignWitnessCollection::store_witness / fetch_witness).

Layout: u32 LE number of advice columns, then column i at byte offset 4 + i * 2^(k+5): 2^k field elements exactly as they
sit in memory (4 x u64 little-endian limbs, Montgomery form) -- the reference writes them through an mmap of the file and
reads them back the same way.  store_witness / fetch_witness here are plain host I/O (no GPU); commit_witness_file hands
the file to the engine, which streams it through pinned staging into the commit pipeline
(plonk/prover.rs:293-299: commit_lagrange_with_bound per advice column)."""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np

from ._lib import check, lib, require_gpu


def store_witness(path: str, advice: Sequence[np.ndarray], k: int) -> None:
    """helpers.rs:920-986 (the file part): advice = columns of shape (2^k, 4) uint64"""
    n = 1 << k
    bundle = k + 5
    with open(path, "wb") as f:
        f.truncate(4 + (len(advice) << bundle))
        f.write(np.uint32(len(advice)).tobytes())
    for i, col in enumerate(advice):
        col = np.ascontiguousarray(col, dtype=np.uint64)
        if col.shape != (n, 4):
            raise ValueError(f"advice column {i} has shape {col.shape}, expected {(n, 4)}")
        mm = np.memmap(path, dtype=np.uint8, mode="r+", offset=4 + (i << bundle), shape=(1 << bundle,))
        mm[:] = col.view(np.uint8).reshape(-1)
        mm.flush()
        del mm


def witness_columns(path: str) -> int:
    with open(path, "rb") as f:
        head = f.read(4)
    if len(head) != 4:
        raise ValueError(f"{path}: no header")
    return int(np.frombuffer(head, dtype="<u4")[0])


def fetch_witness(path: str, k: int) -> List[np.ndarray]:
    """helpers.rs:988-1014: every column copied out of the mapping into its own array"""
    n, bundle = 1 << k, k + 5
    out = []
    for i in range(witness_columns(path)):
        mm = np.memmap(path, dtype=np.uint64, mode="r", offset=4 + (i << bundle), shape=(n, 4))
        out.append(np.array(mm))
        del mm
    return out


def commit_witness_file(params, path: str, first: int = 0, count: Optional[int] = None, max_bits: int = 254,
                        d_keep: int = 0) -> np.ndarray:
    """Commitments (count, 12) of advice columns [first, first + count) of the file against params.g_lagrange, read and
    committed by the engine (b2_commit_witness_file).  d_keep: optional device pointer where the columns stay resident."""
    require_gpu()
    if count is None:
        count = witness_columns(path) - first
    out = np.zeros((max(count, 1), 12), dtype=np.uint64)
    check(lib().b2_commit_witness_file(ctypes.c_uint64(params.g_lagrange.handle), os.fsencode(path), params.k, first, count,
                                       max_bits, ctypes.c_void_p(d_keep or None), ctypes.c_void_p(out.ctypes.data)))
    return out[:count]
