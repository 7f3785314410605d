#!/bin/bash
# round 2, GPU call 29: small MSMs (2^16, 2^18, 2^20) back to back on 1 / 2 / 3 caller streams
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
: > $O/r2_small_msm_streams.jsonl
for lg in 16 18 20; do for v in 1 2 3; do
  python bench.py --logn $lg --steps 60 --warmup 6 --value-streams $v --no-strong --no-quotient --no-ntt --no-proof --no-proof22 --no-cpu > $O/_s.json 2> $O/_s.err
  python -c "
import json; d = json.loads(open('$O/_s.json').read().strip().splitlines()[-1]); print(json.dumps({'logn': $lg, 'streams': $v, 'mpts_s': round(d['value'], 1), 'ms_per_msm': round(d['ms_per_step'], 4), 'e2e_mpts_s': round(d['e2e']['value'], 1), 'parity': d['parity_check']['ok']}))" | tee -a $O/r2_small_msm_streams.jsonl
  tail -c 200 $O/_s.err
done; done
rm -f $O/_s.json $O/_s.err
