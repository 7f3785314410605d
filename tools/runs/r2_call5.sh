#!/bin/bash
# round 2, GPU call 5 (N GPUs): sharded prover with the witness divided over the ranks' PCIe links + NVLink exchange
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=${N:-2}
(free -g | head -2; nproc; nvidia-smi -L; nvidia-smi topo -m | head -12) > $O/r2_box_${N}gpu.txt 2>&1
head -3 $O/r2_box_${N}gpu.txt
run() {  # k, extra flags, tag
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_proof_check.py --circuit zkwasm --k $1 $2 --reps 3 > $O/r2_sharded_${3}_${N}gpu.log 2>&1
  rc=$?; echo "rc=$rc"; grep -h '^{' $O/r2_sharded_${3}_${N}gpu.log | tail -1 | cut -c1-1500
  grep -h "Error\|error\|Traceback" $O/r2_sharded_${3}_${N}gpu.log | head -5
}
for k in ${KS:-18 22}; do
  run $k "--split-quotient" k${k}q
  if [ "$rc" != "0" ]; then echo "stopping: k=$k failed"; tail -30 $O/r2_sharded_k${k}q_${N}gpu.log; break; fi
done
