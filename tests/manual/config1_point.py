#!/usr/bin/env python
"""BASELINE.json configs[0]: the reference's own CPU-runnable case -- BN254 G1 MSM of 2^16 random
points and an Fr NTT at k = 16 through best_multiexp / best_fft -- timed with the C restatement of the reference's rayon path
(oracle/cpu_ref.c: chunk = n / T, one multiexp_serial per thread, ordered fold; arithmetic.rs:465-492, 546-705) on this
box's host cores, next to the engine on the same inputs, AND compared bit for bit (this size the oracle finishes in
milliseconds, so the timed inputs are also a parity case).  It lives under tests/ because it runs oracle/ (as the CPU baseline and as the checker).

    python tests/manual/config1_point.py [--out file.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _lib  # noqa: E402
from halo2_gpu_specific_b200.arithmetic import Srs  # noqa: E402
from oracle import cref  # noqa: E402


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t)
    return min(ts), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    _lib.require_gpu()
    _lib.set_device(0)
    k, n = a.k, 1 << a.k
    cores = os.cpu_count() or 1
    scalars = cref.random_fr_mont(n, 0xB2000001)                      # seed 0xB200_0001
    ks = np.zeros((n, 4), dtype=np.uint64)
    ks[:, 0] = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    bases = cref.g1_mul_gen(ks, threads=cores)                        # [s_i] G
    dom = h2.EvaluationDomain(1, k)

    # ---- MSM
    cpu_s, want = best_of(lambda: cref.best_multiexp(scalars, bases, threads=cores), 3)
    srs = Srs.register(bases)
    h2.best_multiexp(scalars, srs)                                    # plain bases (what a one-off call sees)
    gpu_plain_s, got = best_of(lambda: h2.best_multiexp(scalars, srs), a.reps)
    ok_msm = bool(np.array_equal(got[:8], cref.jac_to_affine(want)[0]))
    srs.precompute()
    h2.best_multiexp(scalars, srs)
    gpu_table_s, got2 = best_of(lambda: h2.best_multiexp(scalars, srs), a.reps)
    ok_msm &= bool(np.array_equal(got2[:8], cref.jac_to_affine(want)[0]))
    kernel_ms = _lib.last_timing()[0]

    # ---- NTT
    x = cref.random_fr_mont(n, 0xB2000011)
    cpu_ntt_s, want_ntt = best_of(lambda: cref.best_fft(x, dom.omega, k, threads=cores), 3)
    buf = x.copy()
    h2.best_fft(buf, dom.omega, k)

    def gpu_fft():
        b = x.copy()
        h2.best_fft(b, dom.omega, k)
        return b
    gpu_ntt_s, got_ntt = best_of(gpu_fft, a.reps)
    ok_ntt = bool(np.array_equal(got_ntt, want_ntt))
    ntt_kernel_ms = _lib.last_timing()[0]

    doc = {"workload": f"BASELINE configs[0]: MSM of 2^{k} points + Fr NTT k = {k} via best_multiexp / best_fft",
           "cpu": {"kind": "port (oracle/cpu_ref.c, restatement of arithmetic.rs)", "cores": cores,
                   "msm_ms": cpu_s * 1e3, "msm_mpts_s": n / cpu_s / 1e6, "ntt_ms": cpu_ntt_s * 1e3, "ntt_melem_s": n / cpu_ntt_s / 1e6},
           "engine": {"msm_host_api_ms_plain_bases": gpu_plain_s * 1e3, "msm_host_api_ms_window_table": gpu_table_s * 1e3,
                      "msm_kernel_ms_window_table": kernel_ms, "msm_mpts_s_host_api": n / gpu_table_s / 1e6,
                      "ntt_host_api_ms": gpu_ntt_s * 1e3, "ntt_kernel_ms": ntt_kernel_ms, "ntt_melem_s_host_api": n / gpu_ntt_s / 1e6,
                      "note": "host API: pageable numpy inputs, H2D + kernels + D2H (+ a host copy of the column for the NTT) per call"},
           "bit_exact_vs_oracle": {"msm": ok_msm, "ntt": ok_ntt}}
    print(json.dumps(doc), flush=True)
    if a.out:
        json.dump(doc, open(a.out, "w"), indent=1)
    srs.free()
    sys.exit(0 if (ok_msm and ok_ntt) else 1)


if __name__ == "__main__":
    main()
