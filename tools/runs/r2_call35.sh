#!/bin/bash
# round 2, GPU call 35: the final build of libb2pcs.so on the GPU once more (smoke: MSM, iNTT, coset extension
# bit-exact, a whole proof accepted by the oracle verifier)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 70 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/r2_smoke_final.log
