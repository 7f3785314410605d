#!/bin/bash
# round 2, GPU call 12: NTT with batched tile loads / inter-pass twiddle loads (4 or 8 rows in flight per thread)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py tests/test_gpu_quotient.py -m gpu -x -q 2>&1 | tail -3
B2PCS_LIB=$PWD/halo2_gpu_specific_b200/variants/libb2pcs_lq8.so python -m pytest tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -2
: > $O/r2_ntt_variants_d.jsonl
KS=18,20,22,24 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_d.jsonl 2>> $O/r2_ntt_variants_d.err
B2PCS_LIB=$PWD/halo2_gpu_specific_b200/variants/libb2pcs_lq8.so KS=18,20,22,24 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_d.jsonl 2>> $O/r2_ntt_variants_d.err
python -c "
import json
for name, l in zip(('4 rows in flight', '8 rows in flight'), open('$O/r2_ntt_variants_d.jsonl')):
    d = json.loads(l); print(name, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
tail -3 $O/r2_ntt_variants_d.err
python tools/sweep.py --ntt-k 22 --msm-logn "" --cols 16 --reps 3 --out $O/r2_sweep_ntt22.json | tail -1 | cut -c1-600
