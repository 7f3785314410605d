#!/bin/bash
# round 2, GPU call 22: e2e pipeline depth 2 vs 3 at N = 1; prover tests (multi-circuit early transforms)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_prover.py tests/test_gpu_msm.py -m gpu -x -q 2>&1 | tail -3
for d in 2 3; do
  python bench.py --steps 20 --warmup 3 --e2e-depth $d --no-strong --no-quotient --no-ntt --no-proof --no-proof22 --no-cpu > $O/_e2e.json 2> $O/_e2e.err
  python -c "
import json; d = json.loads(open('$O/_e2e.json').read().strip().splitlines()[-1]); print('depth $d', 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['e2e']['ms_per_step'], 3), 'blocking', round(d['e2e']['blocking_call']['value'], 1), 'parity', d['parity_check']['ok'])"
  tail -c 200 $O/_e2e.err
done
rm -f $O/_e2e.json $O/_e2e.err
