#!/bin/bash
# round 2, GPU call 21 (8 GPUs): bench.py exactly as the driver launches the N = 8 line
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r2_bench_8gpu.json 2> $O/r2_bench_8gpu.err
echo "bench rc=$?"; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/r2_bench_8gpu.err | tail -5
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_8gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d.get(k) for k in ('value', 'n_gpus', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'parity', d.get('parity_check', {}).get('ok'))
        print('strong', {k: round(v['mpts_per_s']) for k, v in d['strong_scaling']['sizes'].items()})
        print('sharded', json.dumps(d.get('sharded_create_proof'))[:900])
PY
