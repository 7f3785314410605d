#!/bin/bash
# round 2, GPU call 28 (2 GPUs): early transforms of a rank's advice share -- prover tests on one GPU, then the real
# sharded k = 22 proof on two
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_prover.py -m gpu -x -q 2>&1 | tail -3
N=2
run() {
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_proof_check.py --circuit zkwasm --k $1 --split-quotient --reps 3 > $O/r2_sharded_k${1}q_${N}gpu_b.log 2>&1
  rc=$?; echo "rc=$rc"; grep -h '^{' $O/r2_sharded_k${1}q_${N}gpu_b.log | tail -1 | cut -c1-1300
  grep -h "Error\|error\|Traceback" $O/r2_sharded_k${1}q_${N}gpu_b.log | head -5
}
run 18
run 22
