#!/bin/bash
# round 2, GPU call 24 (2 GPUs): the N = 2 line with 1 and 2 caller streams (partial buffers decoupled from the gathers)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for v in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --value-streams $v --no-strong --no-proof > $O/_v.json 2> $O/_v.err
echo "N=2 streams=$v rc=$?"; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/_v.err | tail -5
python -c "
import json; d = json.loads([l for l in open('$O/_v.json') if l.startswith('{')][-1]); print('N=2 streams $v value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'])"
done
rm -f $O/_v.json $O/_v.err
