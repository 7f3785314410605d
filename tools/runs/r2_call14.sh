#!/bin/bash
# round 2, GPU call 14: NTT register budget -- 3 CTAs per SM at 156 registers, 2 at 180 (does ptxas buy ILP with them?)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
V=$PWD/halo2_gpu_specific_b200/variants
: > $O/r2_ntt_variants_f.jsonl
KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_f.jsonl 2>> $O/r2_ntt_variants_f.err
B2PCS_LIB=$V/libb2pcs_regs.so B2_NTT_VARIANT=4 KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_f.jsonl 2>> $O/r2_ntt_variants_f.err
B2PCS_LIB=$V/libb2pcs_regs.so B2_NTT_VARIANT=5 KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_f.jsonl 2>> $O/r2_ntt_variants_f.err
python -c "
import json
for name, l in zip(('default (128 regs, 4 CTAs/SM)', '156 regs, 3 CTAs/SM', '180 regs, 2 CTAs/SM'), open('$O/r2_ntt_variants_f.jsonl')):
    d = json.loads(l); print(name, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
tail -3 $O/r2_ntt_variants_f.err
