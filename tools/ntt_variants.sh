#!/bin/bash
# NTT pass-structure experiment: digit size cap (B2_NTT_MAXM) vs throughput
for k in 20 22 24; do for m in 8 9 10 11 12; do
  echo -n "k=$k maxm=$m: "; B2_NTT_MAXM=$m python tools/sweep.py --ntt-k $k --msm-logn "" --cols 16 --reps 3 --out /dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ntt', round(d['ntt']['melem_s']), 'Melem/s  ext', round(d['coeff_to_extended']['melem_out_s']))"
done; done
