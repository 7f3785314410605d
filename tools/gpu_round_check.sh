#!/bin/bash
# One gpurun call that validates the prover path on a B200: new GPU tests, smoke(), the default bench line,
# then (best effort) the rest of the GPU suite.  Usage: gpurun --timeout 840 -- bash tools/gpu_round_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 330 python -m pytest tests/test_gpu_prover.py -q -s --tb=short -m gpu > gpurun_out/prover_tests.log 2>&1
echo "prover tests rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"
timeout 330 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
timeout 240 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_prover.py > gpurun_out/gpu_suite.log 2>&1
echo "suite rc=$?"
tail -15 gpurun_out/prover_tests.log
tail -3 gpurun_out/smoke.log
tail -3 gpurun_out/gpu_suite.log
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "create_proof", d.get("create_proof"))
except Exception as e:
    print("bench parse failed", e)
P
