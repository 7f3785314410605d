#!/bin/bash
# round 2, GPU call 16: NTT with two stages per shared-memory round trip (4 elements per thread) at 4 and 6 warps per scheduler
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
V=$PWD/halo2_gpu_specific_b200/variants
: > $O/r2_ntt_variants_g.jsonl
KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_g.jsonl 2>> $O/r2_ntt_variants_g.err
B2PCS_LIB=$V/libb2pcs_r2m2.so KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_g.jsonl 2>> $O/r2_ntt_variants_g.err
B2PCS_LIB=$V/libb2pcs_r2m3.so KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_g.jsonl 2>> $O/r2_ntt_variants_g.err
python -c "
import json
for name, l in zip(('default (3 stages per trip, 8 elements per thread, 4 warps/scheduler)', '2 stages per trip, 4 warps/scheduler', '2 stages per trip, 6 warps/scheduler'), open('$O/r2_ntt_variants_g.jsonl')):
    d = json.loads(l); print(name, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
tail -3 $O/r2_ntt_variants_g.err
