#!/usr/bin/env python
"""BASELINE.json configs 2 and 3 on one B200: NTT / iNTT / coset-extend sweep (k = 18..24, 64 columns,
device resident, kernel-only CUDA-event time) and MSM sweep (2^18..2^26; uniform 254-bit, 16-bit
"advice-like" with max_bits = 16, and 50 % zeros).  Writes one JSON document.  No oracle, no torch."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _fr, _lib  # noqa: E402
from halo2_gpu_specific_b200._lib import NttDesc  # noqa: E402
from halo2_gpu_specific_b200.arithmetic import Srs  # noqa: E402


def rand_scalars(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def small_scalars_mont(n, seed, bits):
    """v < 2^bits in Montgomery form: v * R mod r computed with Python ints per distinct value (bits <= 16)"""
    rng = np.random.default_rng(seed)
    v = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
    table = np.stack([_fr.to_mont(int(x)) for x in range(1 << bits)])
    return table[v]


def ntt_sweep(L, ks, cols, reps, out):
    vp = ctypes.c_void_p
    for k in ks:
        n = 1 << k
        dom = h2.EvaluationDomain(5, k)
        ext_k = dom.extended_k
        col = rand_scalars(n, 100 + k)
        # device buffers: input batch (cols x n) and extended output (sub-batched by the library)
        c_here = cols
        while c_here * (n << (ext_k - k)) * 32 > (48 << 30) and c_here > 1:
            c_here //= 2
        d_in, d_out = vp(), vp()
        _lib.check(L.b2_dev_alloc(c_here * n * 32, ctypes.byref(d_in)))
        _lib.check(L.b2_dev_alloc(c_here * (1 << ext_k) * 32, ctypes.byref(d_out)))
        for c in range(c_here):
            _lib.check(L.b2_memcpy_h2d(vp(d_in.value + c * n * 32), _lib.ptr(col), n * 32))

        def run(desc):
            ts = []
            for _ in range(reps + 2):
                _lib.check(L.b2_ntt_exec(ctypes.byref(desc)))
                ts.append(_lib.last_timing()[0])
            return float(np.median(ts[2:]))

        res = {"k": k, "columns": c_here}
        d = NttDesc()
        d.log_n, d.location, d.omega = k, 1, dom.omega.ctypes.data
        d.n_in = d.n_out = d.in_stride = d.out_stride = n
        d.columns, d.in_, d.out = c_here, d_in.value, d_in.value
        ms = run(d)
        res["ntt"] = {"ms": ms, "melem_s": c_here * n / ms / 1e3, "hbm_gbs_algorithmic": 64 * c_here * n / ms / 1e6}
        d.omega, d.divisor = dom.omega_inv.ctypes.data, dom.ifft_divisor.ctypes.data
        ms = run(d)
        res["intt"] = {"ms": ms, "melem_s": c_here * n / ms / 1e3}
        z = np.concatenate([dom.g_coset, dom.g_coset_inv])
        e = NttDesc()
        e.log_n, e.location, e.omega = ext_k, 1, dom.extended_omega.ctypes.data
        e.coset_in = z.ctypes.data
        e.n_in, e.in_stride = n, n
        e.n_out = e.out_stride = 1 << ext_k
        e.columns, e.in_, e.out = c_here, d_in.value, d_out.value
        ms = run(e)
        res["coeff_to_extended"] = {"ext_k": ext_k, "ms": ms, "melem_out_s": c_here * (1 << ext_k) / ms / 1e3}
        zi = np.concatenate([dom.g_coset_inv, dom.g_coset])
        f = NttDesc()
        f.log_n, f.location, f.omega = ext_k, 1, dom.extended_omega_inv.ctypes.data
        f.divisor, f.coset_out = dom.extended_ifft_divisor.ctypes.data, zi.ctypes.data
        f.n_in = f.in_stride = f.out_stride = 1 << ext_k
        f.n_out = n * dom.quotient_poly_degree
        f.columns, f.in_, f.out = 1, d_out.value, d_out.value
        ms = run(f)
        res["extended_to_coeff"] = {"ext_k": ext_k, "ms": ms, "melem_in_s": (1 << ext_k) / ms / 1e3}
        L.b2_dev_free(d_in)
        L.b2_dev_free(d_out)
        print(json.dumps(res), flush=True)
        out["ntt"].append(res)


WINDOW_BITS = 0


def msm_sweep(L, logns, reps, out):
    for lg in logns:
        n = 1 << lg
        t0 = time.time()
        srs = Srs.synthetic(n, 0, 0xB2000003).precompute(WINDOW_BITS)
        setup_s = time.time() - t0
        res = {"logn": lg, "srs_setup_s": setup_s}
        cfg = (ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32())
        L.b2_msm_config(srs.handle, n, 254, *[ctypes.byref(x) for x in cfg])
        res["window_bits"], res["windows"] = cfg[0].value, cfg[1].value
        cases = {"uniform254": (rand_scalars(n, 7), 254)}
        if lg <= 24:
            sm = small_scalars_mont(n, 8, 16)
            cases["u16_maxbits16"] = (sm, 16)
            hz = rand_scalars(n, 9)
            hz[::2] = 0
            cases["half_zero254"] = (hz, 254)
        for name, (sc, bits) in cases.items():
            h = _lib.pinned_empty((n, 4))
            h[:] = sc
            ks, ts = [], []
            for _ in range(reps + 1):
                h2.gpu_multiexp_single_gpu_with_bound(h, srs, bits)
                ks.append(_lib.last_msm_phases()["total"])
                ts.append(_lib.last_timing()[1])
            _lib.pinned_free(h)
            km, tm = float(np.median(ks[1:])), float(np.median(ts[1:]))
            res[name] = {"kernel_ms": km, "e2e_ms": tm, "mpts_s_kernel": n / km / 1e3, "mpts_s_e2e": n / tm / 1e3,
                         "phases": {k_: round(v, 3) for k_, v in _lib.last_msm_phases().items()}}
        srs.free()
        print(json.dumps(res), flush=True)
        out["msm"].append(res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ntt-k", default="18,19,20,21,22,23,24")
    ap.add_argument("--msm-logn", default="18,20,22,24,26")
    ap.add_argument("--cols", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--window-bits", type=int, default=0, help="force the SRS window-table width (0 = library default)")
    a = ap.parse_args()
    global WINDOW_BITS
    WINDOW_BITS = a.window_bits
    _lib.require_gpu()
    L = _lib.lib()
    out = {"ntt": [], "msm": []}
    macs, muls, dfma = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _lib.check(L.b2_imad_probe(ctypes.byref(macs), ctypes.byref(muls)))
    _lib.check(L.b2_dfma_probe(ctypes.byref(dfma)))
    out["probe"] = {"modmul_per_s": muls.value, "wide_mac_per_s": macs.value, "dfma_per_s": dfma.value}
    print(json.dumps(out["probe"]), flush=True)
    if a.ntt_k:
        ntt_sweep(L, [int(x) for x in a.ntt_k.split(",")], a.cols, a.reps, out)
    if a.msm_logn:
        msm_sweep(L, [int(x) for x in a.msm_logn.split(",")], a.reps, out)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
