#!/bin/bash
# round 2, GPU call 15: buckets per reduce thread (B2_MSM_RM) against the reduce kernel's latency chain
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
: > $O/r2_msm_rm.jsonl
for rm in 0 2 4 8 16; do
  B2_MSM_RM=$rm python tools/sweep.py --ntt-k "" --msm-logn 18,20,22 --reps 5 --out $O/_rm.json > /dev/null 2>&1
  python -c "
import json
d = json.load(open('$O/_rm.json'))
out = {'rm': $rm}
for r in d['msm']:
    u = r['uniform254']; out['2^%d' % r['logn']] = {'ms': round(u['kernel_ms'], 3), 'reduce': u['phases']['reduce'], 'final': u['phases']['final'], 'mpts': round(u['mpts_s_kernel'])}
print(json.dumps(out))" | tee -a $O/r2_msm_rm.jsonl
done
rm -f $O/_rm.json
