#!/usr/bin/env python
"""Minimal driver for ncu captures: a few MSMs (2^logn, resident synthetic SRS) and a batch of NTTs.
No torch, no oracle.  Usage: python tools/profile_run.py [--logn 22] [--reps 3] [--ntt-cols 8] [--what msm,ntt]"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _lib  # noqa: E402
from halo2_gpu_specific_b200.arithmetic import Srs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=22)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ntt-cols", type=int, default=8)
    ap.add_argument("--what", default="msm,ntt")
    ap.add_argument("--bits", type=int, default=254)
    ap.add_argument("--precompute", action="store_true")
    a = ap.parse_args()
    _lib.require_gpu()
    n = 1 << a.logn
    rng = np.random.default_rng(1)
    if "msm" in a.what:
        srs = Srs.synthetic(n, 0, 0xB2000003)
        if a.precompute:
            srs.precompute()
        sc = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
        sc[:, 3] &= np.uint64((1 << 60) - 1)
        if a.bits < 254:
            sc[:, 1:] = 0
            sc[:, 0] >>= np.uint64(64 - a.bits)
            # canonical small value v -> Montgomery form would need a multiply; instead keep max_bits=254
        for _ in range(a.reps):
            h2.gpu_multiexp_single_gpu_with_bound(sc, srs, 254)
            print("msm phases (ms):", {k: round(v, 3) for k, v in _lib.last_msm_phases().items()}, flush=True)
    if "ntt" in a.what:
        dom = h2.EvaluationDomain(5, a.logn)
        x = rng.integers(0, 2**64, size=(a.ntt_cols, n, 4), dtype=np.uint64)
        x[:, :, 3] &= np.uint64((1 << 60) - 1)
        for _ in range(a.reps):
            dom.lagrange_to_coeff_batch(x)
            print("ntt kernel/total ms:", _lib.last_timing(), flush=True)


if __name__ == "__main__":
    main()
