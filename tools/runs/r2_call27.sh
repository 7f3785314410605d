#!/bin/bash
# round 2, GPU call 27: fused a*b +- c*d instruction in the quotient kernel -- parity tests, kernel and proof timings with
# and without the fusion (B2_Q_NO_FUSE=1)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_quotient.py tests/test_gpu_grand_product.py tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -3
for f in 0 1; do
  [ $f = 1 ] && export B2_Q_NO_FUSE=1 || unset B2_Q_NO_FUSE
  python bench.py --steps 3 --warmup 3 --no-strong --no-ntt --no-cpu > $O/_q.json 2> $O/_q.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/_q.json').read().strip().splitlines()[-1])
q, p, p18 = d['quotient'], d['create_proof_k22'], d['create_proof']
print('no_fuse=$f', 'quotient', round(q.get('value', 0), 4), 'eval_kernel_ms', round(q.get('eval_kernel_ms', 0), 1), q.get('program'), 'frac', round(q['roofline_int']['frac'], 3))
print('   proof22', round(p.get('value', 0), 4), {k: round(v, 3) for k, v in (p.get('phases_s') or {}).items()}, p.get('h_program'), p.get('error'))
print('   proof18', round(p18.get('value', 0), 5))
PY
done
rm -f $O/_q.json $O/_q.err
