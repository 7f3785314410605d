#!/bin/bash
# What DESIGN.md section 8 lists as the first GPU session of round 2, as commands.
#   1 GPU :  gpurun --timeout 600 -- bash tools/runs/round2_first_run.sh single
#   N GPUs:  gpurun --gpus N --timeout 600 -- bash tools/runs/round2_first_run.sh multi N
mkdir -p gpurun_out
case "$1" in
  single)
    # random circuits through both device engines against the oracle prover's bytes (not yet run on a GPU)
    timeout 200 python tests/manual/gpu_prover_fuzz.py --seeds 40 > gpurun_out/gpu_prover_fuzz.log 2>&1
    echo "gpu fuzz rc=$?"; tail -2 gpurun_out/gpu_prover_fuzz.log
    # the sharded engines on one rank: same bytes, on-device paths exercised
    timeout 120 python tools/sharded_proof_check.py --circuit zkwasm --k 14 > gpurun_out/sharded_1gpu_k14.json 2>&1
    echo "sharded x1 rc=$?"; tail -1 gpurun_out/sharded_1gpu_k14.json
    ;;
  multi)
    N=${2:-2}
    # every rank holds its own copy of the witness and the proving key on the host (about 54 GiB at k = 22): 196 GB of
    # host memory carry two ranks at k = 22, four at k = 21, eight at k = 20 (tools/sharded_proof_check.py refuses otherwise)
    K=22; [ "$N" -ge 4 ] && K=21; [ "$N" -ge 8 ] && K=20
    for extra in "" "--split-quotient"; do
      tag=$( [ -z "$extra" ] && echo commits || echo commits_quotient )
      timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
        --master-port 29517 tools/sharded_proof_check.py --circuit zkwasm --k "$K" $extra \
        > "gpurun_out/sharded_${N}gpu_k${K}_${tag}.log" 2>&1
      echo "sharded x$N $tag rc=$?"; tail -1 "gpurun_out/sharded_${N}gpu_k${K}_${tag}.log"
    done
    ;;
  *)
    echo "usage: $0 single | multi N"; exit 2;;
esac
