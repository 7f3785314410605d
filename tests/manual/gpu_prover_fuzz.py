#!/usr/bin/env python
"""Random circuits (tests/test_prover_fuzz.py's generator) through the DEVICE engines: proof bytes of ResidentEngine and of
the host-API Engine must equal the oracle prover's.  Not part of the pytest suite yet (written after the round's GPU budget
was spent); run it on a B200 box first thing in round 2:

    python tests/manual/gpu_prover_fuzz.py [--seeds 40]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _lib  # noqa: E402
from halo2_gpu_specific_b200 import plonk as HP  # noqa: E402
from oracle import bn254 as o  # noqa: E402
from oracle import prover as PR  # noqa: E402
import test_prover_fuzz as F  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=40)
    a = ap.parse_args()
    _lib.require_gpu()
    _lib.set_device(0)
    oparams = PR.Params(F.K, F.S_TOXIC)
    params = h2.Params(F.K, oparams.g, oparams.g_lagrange)
    bad = []
    for seed in range(a.seeds):
        cs, fixed, advice, instance, mapping = F.build(seed)
        opk = PR.keygen(oparams, cs, fixed, mapping)
        inst = [instance[0][:3]]
        gwc = seed % 2 == 0
        want = PR.create_proof(oparams, opk, advice, inst, HP.SeededRng(seed), use_gwc=gwc)
        hcs = HP.ConstraintSystem.like(cs)
        pk = HP.keygen(params, hcs, np.stack([o.fr_encode(c) for c in fixed]), np.array(mapping, dtype=np.int64),
                       transcript_repr=opk.vk.transcript_repr)
        adv = np.ascontiguousarray(np.stack([o.fr_encode(c) for c in advice]))
        for kind in ("resident", "host_api"):
            eng = (HP.ResidentEngine if kind == "resident" else HP.Engine)(params, pk.vk.domain)
            got = HP.create_proof(params, pk, adv.copy(), inst, HP.SeededRng(seed), engine=eng, use_gwc=gwc)
            eng.free()
            if got != want:
                bad.append((seed, kind))
    print("seeds", a.seeds, "mismatches", bad)
    params.free()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
