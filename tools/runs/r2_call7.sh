#!/bin/bash
# round 2, GPU call 7 (2 GPUs): whole GPU suite incl. the multi-device tests, bench.py as the driver launches it at N = 2
# (own arm and reference arm)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/r2_gpu_suite_c7_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_2gpu.json 2> $O/r2_bench_2gpu.err
echo "bench rc=$?"; tail -c 600 $O/r2_bench_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_2gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print({k: d.get(k) for k in ('metric', 'value', 'n_gpus', 'ms_per_step', 'scaling')}, 'e2e', d['e2e']['value'], 'parity', d.get('parity_check'))
        print('strong', json.dumps(d.get('strong_scaling'))[:800])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2_bench_ref_2gpu.json 2> $O/r2_bench_ref_2gpu.err
echo "ref rc=$?"; cut -c1-600 $O/r2_bench_ref_2gpu.json; tail -c 300 $O/r2_bench_ref_2gpu.err
