#!/usr/bin/env python
"""A/B timing of the NTT kernels (B2_NTT_SHOUP=0/1 in the environment): 64 columns, device resident."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib
from halo2_gpu_specific_b200 import evaluation as E
from halo2_gpu_specific_b200._lib import NttDesc

_lib.require_gpu(); _lib.set_device(0)
L = _lib.lib()
res = {"shoup": os.environ.get("B2_NTT_SHOUP", "1")}
mm, sm = ctypes.c_double(), ctypes.c_double()
L.b2_imad_probe(None, ctypes.byref(mm)); L.b2_shoup_probe(ctypes.byref(sm))
res["montgomery_mul_per_s"] = mm.value; res["shoup_mul_per_s"] = sm.value
for k, cols in [(int(a), 64 if int(a) < 24 else 16) for a in os.environ.get("KS", "18,20,22,24").split(",")]:
    n = 1 << k
    dom = h2.EvaluationDomain(5, k)
    buf = E.DeviceBuffer(cols * n)
    col = np.random.default_rng(k).integers(0, 2**62, size=(n, 4), dtype=np.uint64)
    for c in range(cols): buf.upload(col, c * n)
    d = NttDesc()
    d.log_n, d.location = k, 1
    d.omega, d.divisor = dom.omega.ctypes.data, 0
    d.n_in = d.n_out = d.in_stride = d.out_stride = n
    d.columns = cols
    d.in_ = d.out = buf.ptr
    best = 1e9
    for rep in range(6):
        _lib.check(L.b2_ntt_exec(ctypes.byref(d)))
        km, tm = ctypes.c_double(), ctypes.c_double()
        L.b2_last_timing(ctypes.byref(km), ctypes.byref(tm))
        if rep >= 1: best = min(best, km.value)
    res[f"k{k}x{cols}"] = {"ms": best, "melem_s": cols * n / best / 1e3}
    buf.free()
print(json.dumps(res))
