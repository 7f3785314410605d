#!/bin/bash
# round 2, GPU call 37 (the round's last GPU seconds): the proof circuit in its 17-value gate order, with and without
# the quotient kernel's global slot class
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 14 python tools/gate_order_proof_ab.py --k 18 > gpurun_out/r2_gate_order_proof_ab.json 2> gpurun_out/r2_gate_order_proof_ab.err
echo "rc=$?"; tail -c 1500 gpurun_out/r2_gate_order_proof_ab.json; tail -c 300 gpurun_out/r2_gate_order_proof_ab.err
