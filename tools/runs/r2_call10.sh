#!/bin/bash
# round 2, GPU call 10: lane selection keeps pipelines off lanes with pending caller-stream work; early advice transforms
# measured again; ncu launch list of the bench's MSM step
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_prover.py tests/test_gpu_msm.py tests/test_gpu_ntt.py -m gpu -x -q 2>&1 | tail -5 | tee $O/r2_gpu_c10.log
python bench.py --steps 10 --warmup 3 --no-strong > $O/r2_bench_e.json 2> $O/r2_bench_e.err
tail -c 300 $O/r2_bench_e.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_e.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'], 'ntt', d['ntt']['value'])
p = d['create_proof_k22']
print('proof22', p.get('value'), p.get('phases_s'), p.get('error'))
print('ops', p.get('engine_ops_s_calls'))
print('proof18', d['create_proof'].get('value'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-quotient --no-proof --no-proof22 --no-strong --no-ntt > $O/r2_launches_bench.log 2>&1
grep -c kernel $O/r2_launches_bench_ncu.csv
