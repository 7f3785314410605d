#!/usr/bin/env python
"""Rates of the field-product variants (b2_mul_probe) and the MSM phases with the current build."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import halo2_gpu_specific_b200 as h2
from halo2_gpu_specific_b200 import _lib
from halo2_gpu_specific_b200.arithmetic import Srs
_lib.require_gpu(); _lib.set_device(0)
L = _lib.lib()
res = {}
for kind, name in enumerate(["cios", "shoup", "kara", "sqr_sos", "mul2_add"]):
    if kind in (2, 3):          # generated squaring / Karatsuba: measured slower in round 1, removed
        continue
    v = ctypes.c_double()
    _lib.check(L.b2_mul_probe(kind, ctypes.byref(v)))
    res[name + "_G_per_s"] = round(v.value / 1e9, 2)
n = 1 << 22
srs = Srs.synthetic(n, 0, 0xB2000003).precompute()
sc = np.random.default_rng(1).integers(0, 2**64, size=(n, 4), dtype=np.uint64)
sc[:, 3] &= np.uint64((1 << 60) - 1)
for _ in range(3):
    h2.gpu_multiexp_single_gpu_with_bound(sc, srs, 254)
res["msm_phases_ms"] = {k: round(v, 3) for k, v in _lib.last_msm_phases().items()}
print(json.dumps(res))
