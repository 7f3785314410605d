#!/bin/bash
# round 2, GPU call 25 (2 GPUs): block gather -- N = 1 (no collective), N = 2 with the gather per block and per step
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python bench.py --steps 20 --warmup 3 --no-strong --no-quotient --no-ntt --no-proof --no-proof22 --no-cpu > $O/_v.json 2> $O/_v.err
python -c "
import json; d = json.loads(open('$O/_v.json').read().strip().splitlines()[-1]); print('N=1 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'], 'launches', d['gpu_launches'])"
tail -c 300 $O/_v.err
for g in 0 1; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --gather-every $g --no-strong --no-proof > $O/_v.json 2> $O/_v.err
echo "N=2 gather-every=$g rc=$?"; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/_v.err | tail -5
python -c "
import json; d = json.loads([l for l in open('$O/_v.json') if l.startswith('{')][-1]); print('N=2 gather-every $g value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'])"
done
rm -f $O/_v.json $O/_v.err
