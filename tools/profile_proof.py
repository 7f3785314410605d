#!/usr/bin/env python
"""Driver for an ncu launch list of the prover path: builds the SRS on the device, runs keygen and two proofs of the
benches/plonk.rs circuit (the second one is the steady state: proving key and buffers resident).  Prints the number of
library launches before the second proof so that the list can be cut there.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_proof.csv \\
        python tools/profile_proof.py --k 18
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import halo2_gpu_specific_b200 as h2  # noqa: E402
from halo2_gpu_specific_b200 import _lib  # noqa: E402
from halo2_gpu_specific_b200 import plonk as HP  # noqa: E402
import plonk_bench_circuit as bc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=18)
    ap.add_argument("--shplonk", action="store_true")
    a = ap.parse_args()
    _lib.require_gpu()
    _lib.set_device(0)
    params = h2.Params.unsafe_setup(a.k, 0x2B200B200B200B200B200B200B200B2001)
    cs = HP.ConstraintSystem(**bc.constraint_system_args())
    fixed, advice, mapping = bc.build(a.k)
    pk = HP.keygen(params, cs, fixed, mapping)
    eng = HP.ResidentEngine(params, pk.vk.domain)
    L = _lib.lib()
    HP.create_proof(params, pk, advice.copy(), [], HP.SeededRng(0), engine=eng, use_gwc=not a.shplonk)
    before = int(L.b2_launch_count(0))
    tm = {}
    HP.create_proof(params, pk, advice.copy(), [], HP.SeededRng(1), engine=eng, use_gwc=not a.shplonk, timings=tm)
    after = int(L.b2_launch_count(0))
    print(f"launches_before_second_proof={before} launches_of_second_proof={after - before} phases={tm}", flush=True)
    eng.free()
    params.free()


if __name__ == "__main__":
    main()
