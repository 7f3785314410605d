#!/bin/bash
# round 2, GPU call 36: the marginal cases of the slot-class policy -- 8 and 10 live values (all shared: 6 / 5 CTAs per
# SM; hybrid: 7 + 1 / 7 + 3 with 7 CTAs per SM)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
for ll in 6 8; do
  timeout 30 python tools/quotient_bench.py --k 17 --long-lived $ll --reps 5 --json $O/r2_quotient_slot_classes_ll$ll.json > /dev/null 2>&1
  python - <<PY
import json
try:
    d = json.load(open('$O/r2_quotient_slot_classes_ll$ll.json'))
    a = d['ab_global_slot_class']
    print($ll, d['program']['n_slots_shared'], d['program']['n_slots_global'], d['kernel_ms'], a['all_slots_shared']['kernel_ms'], a['results_equal'])
except Exception as e:
    print('no result', e)
PY
done
