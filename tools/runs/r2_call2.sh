#!/bin/bash
# round 2, GPU call 2: pipe probes (+ ncu pipe counters), NTT kernel variants (parity + timing), ncu --set full captures
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
python tools/pipe_probe.py > $O/r2_pipe_probe.json 2> $O/r2_pipe_probe.err
cat $O/r2_pipe_probe.json
ncu --query-metrics 2>/dev/null | grep -iE "pipe_fma|pipe_alu|pipe_imad|inst_executed_pipe" > $O/r2_ncu_metric_names.txt
M=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__inst_executed.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:probe --csv --log-file $O/r2_ncu_pipe_probe.csv python tools/pipe_probe.py > /dev/null 2> $O/r2_ncu_pipe_probe.err
tail -3 $O/r2_ncu_pipe_probe.csv | cut -c1-300
: > $O/r2_ntt_variants.jsonl
for v in 0 1 2 3 4; do
  B2_NTT_VARIANT=$v KS=18,20,22,24 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants.jsonl 2>> $O/r2_ntt_variants.err
done
cat $O/r2_ntt_variants.jsonl | python -c "
import sys, json
for i, l in enumerate(sys.stdin):
    d = json.loads(l); print('variant', i, {k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
for v in 1 2 4; do
  echo "== parity, variant $v"; B2_NTT_VARIANT=$v timeout 900 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_poly.py -x -q 2>&1 | tail -3
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:msm_accumulate -s 1 -c 1 -o $O/r2_prof_acc python tools/profile_run.py --reps 2 --what msm --precompute > $O/r2_prof_acc.log 2>&1
$NCU -k regex:ntt_pass -c 2 -o $O/r2_prof_ntt python tools/profile_run.py --reps 1 --ntt-cols 8 --what ntt > $O/r2_prof_ntt.log 2>&1
B2_NTT_VARIANT=2 $NCU -k regex:ntt_pass -c 2 -o $O/r2_prof_ntt_v2 python tools/profile_run.py --reps 1 --ntt-cols 8 --what ntt > $O/r2_prof_ntt_v2.log 2>&1
$NCU -k regex:quotient_eval -c 1 -o $O/r2_prof_quot python tools/quotient_bench.py --k 20 --reps 1 > $O/r2_prof_quot.log 2>&1
ls -la $O/*.ncu-rep
