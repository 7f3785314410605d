#!/bin/bash
# round 2, GPU call 34 (2 GPUs): bench.py as the driver launches it at N = 2, reduced to the MSM line, e2e, parity check
# and the sharded k = 18 proof (range commits by the cost model, ChaCha20 generator, synchronized rng)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-ntt --no-cpu --no-quotient --no-proof22 --no-strong \
  > $O/r2_bench_2gpu_final.json 2> $O/r2_bench_2gpu_final.err
echo "bench N=2 rc=$?"; tail -c 400 $O/r2_bench_2gpu_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'unit', 'n_gpus', 'ms_per_step', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'parity', d['parity_check'])
print('sharded proof', d.get('sharded_create_proof'))
PY
