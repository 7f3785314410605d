"""Regenerates tests/golden/fixtures.npz from the Python big-int oracle (oracle/bn254.py).

The reference (Rust, un-vendored arithmetic, no toolchain here) cannot produce vectors, and
its own tests hold none; these fixtures freeze the oracle's outputs on fixed seeds so that a
later change to either oracle or the CUDA path is caught.  Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bn254 as o  # noqa: E402


def main():
    out = {}
    # NTT k=10 forward, iNTT, coset extend k=6 -> 8, extended_to_coeff
    k = 10
    x = o.random_fr(1 << k, 0xB2000002)
    dom = o.EvaluationDomain(5, k)
    y = list(x)
    o.best_fft(y, dom.omega, k)
    out["ntt_k10_in"] = o.fr_encode(x)
    out["ntt_k10_omega"] = o.fr_encode([dom.omega])[0]
    out["ntt_k10_out"] = o.fr_encode(y)
    out["intt_k10_out"] = o.fr_encode(dom.lagrange_to_coeff(x))
    dom6 = o.EvaluationDomain(5, 6)
    c = o.random_fr(64, 0xB2000012)
    ext = dom6.coeff_to_extended(c)
    out["ext_k6_in"] = o.fr_encode(c)
    out["ext_k6_out"] = o.fr_encode(ext)
    out["ext_k6_back"] = o.fr_encode(dom6.extended_to_coeff(ext))
    # MSM 2^9 over bases [h_i]G
    n = 512
    hs = o.random_fr(n, 0xB2000013)
    bases = [o.g1_mul(o.G1_GEN, h) for h in hs]
    sc = o.random_fr(n, 0xB2000003)
    res = o.best_multiexp(sc, bases, 8)
    out["msm_n512_scalars"] = o.fr_encode(sc)
    out["msm_n512_bases"] = o.g1_affine_encode(bases)
    out["msm_n512_out"] = o.g1_jacobian_encode(res)
    # Params K=6 (test_commit_lagrange shape): g, g_lagrange, commitment of a[i] = i
    p = o.Params(6, 0x1234567890ABCDEF1234567890ABCDEF % o.R_MOD)
    a = list(range(64))
    out["params_k6_g"] = o.g1_affine_encode(p.g)
    out["params_k6_g_lagrange"] = o.g1_affine_encode(p.g_lagrange)
    out["params_k6_commit_lagrange"] = o.g1_jacobian_encode(p.commit_lagrange(a))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz"), **out)
    print("wrote fixtures.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
