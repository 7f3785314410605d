#!/bin/bash
# One gpurun call that validates the whole tree on a B200: the GPU test suite, smoke(), the default bench line and
# the real zkWasm-shaped k=22 proof.  Usage: gpurun --timeout 840 -- bash tools/gpu_round_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 360 python -m pytest tests -q -x -m gpu > gpurun_out/gpu_suite.log 2>&1
echo "suite rc=$?"
tail -4 gpurun_out/gpu_suite.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
timeout 150 python tests/manual/prove_zkwasm_shape.py --k 22 --reps 2 --out gpurun_out/zkwasm_shape_k22.json > gpurun_out/zk22.log 2>&1
echo "zk22 rc=$?"
tail -c 1800 gpurun_out/zk22.log
timeout 270 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
    cp = d.get("create_proof", {})
    print("bench value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "ntt", d.get("ntt", {}).get("value"),
          "quotient", d.get("quotient", {}).get("value"), "create_proof", cp.get("value"), cp.get("error"))
except Exception as e:
    print("bench parse failed", e)
P
tail -3 gpurun_out/bench_final.err
