#!/bin/bash
# round 2, GPU call 32 (2 GPUs): the real prover divided over two ranks with EVERY full-width block committed by point
# range (instance, random polynomial, h pieces, multiopen witnesses, and -- forced -- the z columns' block too is left
# column-parallel because it goes through commit_lagrange_and_ifft); bytes of both ranks against the single-GPU proof
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  tools/sharded_proof_check.py --circuit zkwasm --k ${K:-18} --split-quotient --reps 2 --range-shard on \
  > $O/r2_sharded_range_on_2gpu.log 2>&1
echo "range on rc=$?"; tail -1 $O/r2_sharded_range_on_2gpu.log | cut -c1-1500
