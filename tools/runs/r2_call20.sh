#!/bin/bash
# round 2, GPU call 20: lane count against the k = 22 proof (the early advice transforms occupy one lane)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for l in 3 4 6; do
  B2_LANES=$l python bench.py --steps 3 --warmup 3 --no-strong --no-quotient --no-ntt --no-proof --no-cpu > $O/_lanes.json 2> $O/_lanes.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/_lanes.json').read().strip().splitlines()[-1])
p = d['create_proof_k22']
print('B2_LANES=$l', 'proof22', round(p.get('value', 0), 4), {k: round(v, 3) for k, v in (p.get('phases_s') or {}).items()}, p.get('error'), 'e2e', round(d['e2e']['value']), 'conc', round(d['e2e_concurrent']['value']))
PY
done
rm -f $O/_lanes.json $O/_lanes.err
