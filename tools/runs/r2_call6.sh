#!/bin/bash
# round 2, GPU call 6: full GPU suite on the current tree, two-pipe probe (fp64 next to IMAD.WIDE?), batched-affine probe,
# pipe probes, default bench
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/r2_gpu_suite_c6.log
python tools/pipe_probe.py > $O/r2_pipe_probe.json 2> $O/r2_pipe_probe.err; cat $O/r2_pipe_probe.json
python tools/two_pipe_probe.py > $O/r2_two_pipe_probe.json 2> $O/r2_two_pipe_probe.err; cat $O/r2_two_pipe_probe.json; tail -3 $O/r2_two_pipe_probe.err
python bench.py --steps 10 --warmup 3 > $O/r2_bench_c.json 2> $O/r2_bench_c.err
tail -c 300 $O/r2_bench_c.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['ok'])
print('roofline', json.dumps(d['roofline'])[:1500])
print('ntt', d['ntt']['value'], 'proof22', d['create_proof_k22'].get('value'))
PY
