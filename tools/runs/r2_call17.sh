#!/bin/bash
# round 2, GPU call 17 (2 GPUs): bench.py at N = 2 with the sharded real proof inside (driver-visible multi-GPU prover parity)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu_b.json 2> $O/r2_bench_2gpu_b.err
echo "bench rc=$?"; tail -c 800 $O/r2_bench_2gpu_b.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_2gpu_b.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'parity', d.get('parity_check', {}).get('ok'))
        print('sharded', json.dumps(d.get('sharded_create_proof'))[:1500])
PY
python -m pytest tests/test_gpu_multidevice.py -m gpu -q 2>&1 | tail -3
