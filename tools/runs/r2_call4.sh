#!/bin/bash
# round 2, GPU call 4 (2 GPUs): the real sharded prover (commitments by column, evaluate_h by rows) on a zkWasm-shaped circuit
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=${N:-2}
(free -g | head -2; nproc; nvidia-smi -L) | tee $O/r2_box_${N}gpu.txt
run() {  # k, extra flags, tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_proof_check.py --circuit zkwasm --k $1 $2 --reps 2 > $O/r2_sharded_${3}_${N}gpu.log 2>&1
  echo "rc=$?"; grep -h '^{' $O/r2_sharded_${3}_${N}gpu.log | tail -1 | cut -c1-1500
  grep -h "Error\|error\|Traceback" $O/r2_sharded_${3}_${N}gpu.log | head -5
}
run 18 "--split-quotient" k18q
run 20 "--split-quotient" k20q
run 22 "--split-quotient" k22q
run 22 "" k22c
