#!/bin/bash
# round 2, GPU call 26: compute-sanitizer memcheck over the tests that touch the code added in round 2
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $O/r2_memcheck.log \
  python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py -m gpu -q -x \
  -k "groups or caller_stream or affine_batch or early_advice or golden_fixture or kat_30G or sharded_engine_on_one_rank or multiplicit" 2>&1 | tail -4
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|Error" $O/r2_memcheck.log | head -10
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $O/r2_racecheck.log \
  python -m pytest tests/test_gpu_ntt.py -m gpu -q -x -k "kat_ntt4 or golden" 2>&1 | tail -3
grep -E "RACECHECK SUMMARY|hazard" $O/r2_racecheck.log | head -5
