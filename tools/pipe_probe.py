#!/usr/bin/env python
"""Integer-pipe denominators of the roofline, measured on the device (no torch, no oracle):
raw IMAD.WIDE.U32 and IMAD rates (b2_pipe_probe: 8 independent chains, no carries), the Montgomery product and the Shoup
constant product (b2_imad_probe / b2_shoup_probe), and the fp64 FMA rate.  Prints one JSON object."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo2_gpu_specific_b200 import _lib  # noqa: E402


def measure():
    _lib.require_gpu()
    _lib.set_device(0)
    L = _lib.lib()
    v = ctypes.c_double()
    out = {}
    for kind, name in ((0, "imad_wide_u32_per_s"), (1, "imad_u32_per_s")):
        _lib.check(L.b2_pipe_probe(kind, ctypes.byref(v)))
        out[name] = v.value
    macs, muls = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.b2_imad_probe(ctypes.byref(macs), ctypes.byref(muls)))
    out["montgomery_products_per_s"] = muls.value
    _lib.check(L.b2_shoup_probe(ctypes.byref(v)))
    out["shoup_products_per_s"] = v.value
    _lib.check(L.b2_dfma_probe(ctypes.byref(v)))
    out["dfma_per_s"] = v.value
    sms, ghz = 148, 1.965
    out["per_clk_per_sm_at_1965MHz"] = {k: out[k] / (sms * ghz * 1e9) for k in ("imad_wide_u32_per_s", "imad_u32_per_s")}
    out["nominal_imad_per_s"] = sms * 64 * ghz * 1e9
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
