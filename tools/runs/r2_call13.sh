#!/bin/bash
# round 2, GPU call 13: NTT A/B -- batched tile loads only, batched inter-pass twiddle loads only, two butterflies per
# trip in the lone stage and the cluster stage
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
V=$PWD/halo2_gpu_specific_b200/variants
B2PCS_LIB=$V/libb2pcs_bBFLY.so python -m pytest tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -2
: > $O/r2_ntt_variants_e.jsonl
KS=20,22,24 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_e.jsonl 2>> $O/r2_ntt_variants_e.err
for v in bLOAD bSTORE bBFLY; do
  B2PCS_LIB=$V/libb2pcs_$v.so KS=20,22,24 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_e.jsonl 2>> $O/r2_ntt_variants_e.err
done
python -c "
import json
for name, l in zip(('default', 'batched tile loads', 'batched inter-pass twiddles', 'two butterflies per trip'), open('$O/r2_ntt_variants_e.jsonl')):
    d = json.loads(l); print(name, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
tail -3 $O/r2_ntt_variants_e.err
