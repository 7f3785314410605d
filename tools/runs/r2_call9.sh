#!/bin/bash
# round 2, GPU call 9: prover tests with the early advice transforms, k = 22 proof with and without them, default bench
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -5 | tee $O/r2_gpu_prover_c9.log
python bench.py --steps 10 --warmup 3 > $O/r2_bench_d.json 2> $O/r2_bench_d.err
tail -c 300 $O/r2_bench_d.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'], 'ntt', d['ntt']['value'])
p = d['create_proof_k22']
print('proof22', p.get('value'), p.get('phases_s'), p.get('error'))
print('ops', p.get('engine_ops_s_calls'))
print('proof18', d['create_proof'].get('value'))
PY
