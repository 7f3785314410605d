#!/bin/bash
# MSM window-width experiment: kernel ms for uniform 254-bit scalars vs the SRS table's window bits
for spec in "16:12 13 14 15 16" "18:14 15 16 17 18" "20:16 17 18 19 20" "22:19 20 21"; do
  lg=${spec%%:*}; for c in ${spec#*:}; do
    echo -n "logn=$lg c=$c: "; python tools/sweep.py --ntt-k "" --msm-logn $lg --reps 3 --window-bits $c --out /dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); v=d['uniform254']; print(d['windows'], round(v['kernel_ms'],3), v['phases'])"
  done; done
