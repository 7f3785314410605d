#!/bin/bash
# round 2, GPU call 18: msm_accumulate DRAM traffic against the prefetch depth (0 / 1 / 2 sectors of the next point)
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
V=$PWD/halo2_gpu_specific_b200/variants
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
for v in default pf1 pf0; do
  lib=$PWD/halo2_gpu_specific_b200/libb2pcs.so; [ $v != default ] && lib=$V/libb2pcs_$v.so
  B2PCS_LIB=$lib ncu --metrics $M --clock-control none -k regex:msm_accumulate -s 1 -c 1 --csv --log-file $O/r2_ncu_acc_$v.csv python tools/profile_run.py --reps 2 --what msm --precompute > /dev/null 2>&1
  echo "== $v"; grep -E "dram__bytes|gpu__time|hit_rate|fmaheavy|srcunit" $O/r2_ncu_acc_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
  B2PCS_LIB=$lib python tools/sweep.py --ntt-k "" --msm-logn 22 --reps 5 --out $O/_pf.json > /dev/null 2>&1
  python -c "
import json; d = json.load(open('$O/_pf.json'))['msm'][0]['uniform254']; print('events: total', round(d['kernel_ms'], 3), 'accumulate', d['phases']['accumulate'])"
done
rm -f $O/_pf.json
