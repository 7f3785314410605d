#!/bin/bash
# round 2, GPU call 19: L2 fetch granularity hint (cudaLimitMaxL2FetchGranularity) against the 2x DRAM over-fetch of the
# MSM's 64-byte gathers; NTT throughput under the same setting
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for g in 128 64 32; do
  echo "== B2_L2_FETCH=$g"
  B2_L2_FETCH=$g ncu --metrics $M --clock-control none -k regex:msm_accumulate -s 1 -c 1 --csv --log-file $O/r2_ncu_acc_l2f$g.csv python tools/profile_run.py --reps 2 --what msm --precompute > /dev/null 2>&1
  grep -E "dram__bytes|gpu__time" $O/r2_ncu_acc_l2f$g.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  B2_L2_FETCH=$g python tools/sweep.py --ntt-k 22 --msm-logn 22 --cols 64 --reps 4 --out $O/_l2.json > /dev/null 2>&1
  python -c "
import json; d = json.load(open('$O/_l2.json')); m = d['msm'][0]['uniform254']; n = d['ntt'][0]
print('events: msm total', round(m['kernel_ms'], 3), 'accumulate', m['phases']['accumulate'], 'scatter', m['phases']['scatter'], '| ntt', round(n['ntt']['melem_s']), 'ext', round(n['coeff_to_extended']['melem_out_s']))"
done
rm -f $O/_l2.json
