#!/usr/bin/env python
"""Two design measurements for the MSM (no torch, no oracle); prints one JSON object.

1. b2_mixed_probe: does the fp64 FMA pipe of B200 run NEXT TO the wide-integer multiplier?  For several splits of the
   8 warps of a block into integer warps and fp64 warps: time of the integer warps alone, of the fp64 warps alone and of
   both together, with iteration counts chosen so that the two solo times are about equal.  `overlap` = (T_int + T_f64 -
   T_both) / min(T_int, T_f64): 1 = perfectly concurrent, 0 = serialised.
2. b2_affine_batch_probe: batched-affine additions (6 products + a shared inversion) against the XYZZ mixed add
   (9.44 product-equivalents) on the same points and the same per-thread memory pattern.
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo2_gpu_specific_b200 import _lib  # noqa: E402


def mixed(L, imask, dmask, kind, it_i, it_d):
    ms = ctypes.c_double()
    _lib.check(L.b2_mixed_probe(imask, dmask, kind, it_i, it_d, ctypes.byref(ms)))
    return ms.value


def main():
    _lib.require_gpu()
    _lib.set_device(0)
    L = _lib.lib()
    out = {"mixed": [], "affine_batch": []}
    if "--affine-only" in sys.argv:          # one configuration, for an ncu capture of the two kernels
        b, x = ctypes.c_double(), ctypes.c_double()
        _lib.check(L.b2_affine_batch_probe(1 << 24, 256, ctypes.byref(b), ctypes.byref(x), None, None, None, 0))
        print(json.dumps({"batch_ms": b.value, "xyzz_ms": x.value}))
        return
    for kind, kname in ((0, "montgomery_products"), (1, "raw_imad_wide")):
        for imask, dmask, label in ((0x0f, 0xf0, "4 int + 4 fp64 warps per block"),
                                    (0x3f, 0xc0, "6 int + 2 fp64"),
                                    (0xff, 0x00, "8 int (reference point)")):
            it_i = 3000
            t_i = mixed(L, imask, 0, kind, it_i, 0)
            rec = {"int_kind": kname, "split": label, "iters_int": it_i, "t_int_ms": t_i}
            if dmask:
                # calibrate the fp64 iteration count so that its solo time matches the integer warps' solo time
                t_probe = mixed(L, 0, dmask, kind, 0, 2000)
                it_d = max(1, int(2000 * t_i / t_probe))
                t_d = mixed(L, 0, dmask, kind, 0, it_d)
                t_b = mixed(L, imask, dmask, kind, it_i, it_d)
                rec.update({"iters_f64": it_d, "t_f64_ms": t_d, "t_both_ms": t_b,
                            "overlap": (t_i + t_d - t_b) / min(t_i, t_d),
                            "dfma_per_s_solo": bin(dmask).count("1") * 32 * 2 * 148 * it_d * 64 / (t_d * 1e-3),
                            "dfma_per_s_mixed_if_concurrent": bin(dmask).count("1") * 32 * 2 * 148 * it_d * 64 / (t_b * 1e-3)})
            nwarps = bin(imask).count("1")
            rec["int_products_per_s_solo"] = nwarps * 32 * 2 * 148 * it_i * 2 / (t_i * 1e-3)
            out["mixed"].append(rec)
    n = 1 << 24
    for B in (64, 256, 512, 1024):
        b, x = ctypes.c_double(), ctypes.c_double()
        _lib.check(L.b2_affine_batch_probe(n, B, ctypes.byref(b), ctypes.byref(x), None, None, None, 0))
        out["affine_batch"].append({
            "pairs": n, "B": B, "threads": n // B, "batch_ms": b.value, "xyzz_ms": x.value,
            "batch_adds_per_s": n / (b.value * 1e-3), "xyzz_adds_per_s": 2 * n / (x.value * 1e-3),
            "per_add_speedup": (x.value / 2) / b.value})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
