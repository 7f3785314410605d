#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
: > $O/r2_lane_rotation.jsonl
python tools/lane_rotation_check.py >> $O/r2_lane_rotation.jsonl 2>&1
B2_LANES=1 python tools/lane_rotation_check.py >> $O/r2_lane_rotation.jsonl 2>&1
B2_LANE_NO_AFFINITY=1 python tools/lane_rotation_check.py >> $O/r2_lane_rotation.jsonl 2>&1
cat $O/r2_lane_rotation.jsonl
python bench.py --steps 10 --warmup 3 --no-strong --no-quotient > $O/r2_bench_f.json 2> $O/r2_bench_f.err
tail -c 300 $O/r2_bench_f.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_f.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'], 'ntt', d['ntt']['value'])
p = d['create_proof_k22']
print('proof22', p.get('value'), p.get('phases_s'), p.get('error'))
print('proof18', d['create_proof'].get('value'))
PY
