#!/bin/bash
# round 2, GPU call 8: NTT A/B (default / called Shoup product / single-CTA 2^11 tiles), ncu pipe counters of the raw
# probes, ncu --set full of the batched-affine probe kernels
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
: > $O/r2_ntt_variants_c.jsonl
KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_c.jsonl 2>> $O/r2_ntt_variants_c.err
B2PCS_LIB=$PWD/halo2_gpu_specific_b200/variants/libb2pcs_noinline.so KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_c.jsonl 2>> $O/r2_ntt_variants_c.err
B2_NTT_CLUSTER=0 KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_c.jsonl 2>> $O/r2_ntt_variants_c.err
B2_NTT_VARIANT=2 KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_c.jsonl 2>> $O/r2_ntt_variants_c.err
python -c "
import json
for name, l in zip(('default (lazy)', 'called Shoup product', 'single CTA per 2^11 tile', 'lazy + TMA-staged twiddles'), open('$O/r2_ntt_variants_c.jsonl')):
    d = json.loads(l); print(name, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
tail -3 $O/r2_ntt_variants_c.err
M=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:probe --csv --log-file $O/r2_ncu_pipe_probe.csv python tools/pipe_probe.py > /dev/null 2> $O/r2_ncu_pipe_probe.err
grep -c probe $O/r2_ncu_pipe_probe.csv
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:affine_batch_probe\|xyzz_pair_probe -c 2 -o $O/r2_prof_affine python tools/two_pipe_probe.py --affine-only > $O/r2_prof_affine.log 2>&1
ncu -i $O/r2_prof_affine.ncu-rep --page raw --csv > $O/r2_prof_affine.raw.csv 2>/dev/null
rm -f $O/r2_prof_affine.ncu-rep
ls -la $O | tail -6
