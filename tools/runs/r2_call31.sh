#!/bin/bash
# round 2, GPU call 31: A/B of the quotient kernel's global slot class on programs with long-lived values, then the
# round-end check (whole GPU suite incl. the new range-commit and slot-class tests, smoke, default bench line)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for ll in 16 32; do
  timeout 120 python tools/quotient_bench.py --k 18 --long-lived $ll --reps 5 --json $O/r2_quotient_slot_classes_ll$ll.json > $O/r2_quotient_slot_classes_ll$ll.log 2>&1
  echo "slot classes ll=$ll rc=$?"; python - <<PY
import json
try:
    d = json.load(open('$O/r2_quotient_slot_classes_ll$ll.json'))
    print(d['program'], d['kernel_ms'], d.get('ab_global_slot_class'))
except Exception as e:
    print('no result', e); print(open('$O/r2_quotient_slot_classes_ll$ll.log').read()[-1500:])
PY
done
bash tools/gpu_round_check.sh
