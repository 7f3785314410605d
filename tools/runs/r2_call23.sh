#!/bin/bash
# round 2, GPU call 23 (2 GPUs): device-resident loop on 1 / 2 / 3 caller streams at N = 1, and the N = 2 line with 2
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for v in 1 2 3; do
  python bench.py --steps 20 --warmup 3 --value-streams $v --no-strong --no-quotient --no-ntt --no-proof --no-proof22 --no-cpu > $O/_v.json 2> $O/_v.err
  python -c "
import json; d = json.loads(open('$O/_v.json').read().strip().splitlines()[-1]); print('streams $v', 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check']['ok'], 'launches', d['gpu_launches'])"
  tail -c 300 $O/_v.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-strong --no-proof > $O/_v.json 2> $O/_v.err
echo "N=2 rc=$?"; grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" $O/_v.err | tail -5
python -c "
import json; d = json.loads([l for l in open('$O/_v.json') if l.startswith('{')][-1]); print('N=2 value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'parity', d['parity_check'])"
rm -f $O/_v.json $O/_v.err
