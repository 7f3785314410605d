#!/bin/bash
# What the driver runs at round end, in one call: the whole GPU suite, smoke(), the default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $O/r2_gpu_suite_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/r2_smoke_final.log
python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err
echo "bench rc=$?"; tail -c 300 $O/r2_bench_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'gpu_launches')})
print('e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'], 'roofline', d['roofline']['frac'], d['roofline']['frac_vs'].get('raw_imad_wide_probe'))
print('ntt', d['ntt']['value'], 'proof18', d['create_proof'].get('value'), 'proof22', d['create_proof_k22'].get('value'), 'cpu', d['cpu_baseline']['value'])
PY
