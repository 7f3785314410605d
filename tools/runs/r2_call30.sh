#!/bin/bash
# round 2, GPU call 30: ncu --set full of the quotient kernel with the fused two-product instruction (and without)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
for f in fused plain; do
  [ $f = plain ] && export B2_Q_NO_FUSE=1 || unset B2_Q_NO_FUSE
  $NCU -k regex:quotient_eval -c 1 -o $O/r2_prof_quot_$f python tools/quotient_bench.py --k 20 --reps 1 > $O/r2_prof_quot_$f.log 2>&1
  ncu -i $O/r2_prof_quot_$f.ncu-rep --page raw --csv > $O/r2_prof_quot_$f.raw.csv 2>/dev/null
  rm -f $O/r2_prof_quot_$f.ncu-rep
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/r2_prof_quot_$f.raw.csv')))
d = dict(zip(rows[0], rows[2]))
print('$f', {k: d.get(k) for k in ('gpu__time_duration.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'smsp__inst_executed.sum')})
PY
done
