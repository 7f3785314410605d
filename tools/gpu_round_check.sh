#!/bin/bash
# One gpurun call that validates the whole tree on a B200: the GPU test suite, smoke() and the default bench line
# (which includes the real proofs).  Usage: gpurun --timeout 840 -- bash tools/gpu_round_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests -q -x -m gpu > gpurun_out/gpu_suite.log 2>&1
echo "suite rc=$?"
tail -4 gpurun_out/gpu_suite.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
timeout 420 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    cp, c22 = d.get("create_proof", {}), d.get("create_proof_k22", {})
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "ntt", d.get("ntt", {}).get("value"),
          "quotient", d.get("quotient", {}).get("value"), "create_proof", cp.get("value"), cp.get("error"))
    print("k22", c22.get("value"), c22.get("all_s"), c22.get("phases_s"), c22.get("error"), c22.get("trace"))
except Exception as e:
    print("bench parse failed", e)
P
tail -3 gpurun_out/bench_final.err
