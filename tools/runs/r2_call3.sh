#!/bin/bash
# round 2, GPU call 3: full GPU suite, NTT register variants, ncu --set full captures exported to CSV on the box
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
(free -g | head -2; nproc; df -h /dev/shm | tail -1) | tee $O/r2_box.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
: > $O/r2_ntt_variants_b.jsonl
for v in 0 1 4 5; do
  B2_NTT_VARIANT=$v KS=20,22 python tests/manual/ntt_ab.py >> $O/r2_ntt_variants_b.jsonl 2>> $O/r2_ntt_variants_b.err
done
python -c "
import sys, json
for l in open('$O/r2_ntt_variants_b.jsonl'):
    d = json.loads(l); print({k: round(v['melem_s']) for k, v in d.items() if k.startswith('k')})"
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, extra ncu args, command...
  name=$1; re=$2; extra=$3; shift 3
  $NCU -k regex:$re $extra -o $O/$name "$@" > $O/$name.log 2>&1
  ncu -i $O/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i $O/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/$name.source.csv.gz
  rm -f $O/$name.ncu-rep
}
cap r2_prof_acc msm_accumulate "-s 1 -c 1" python tools/profile_run.py --reps 2 --what msm --precompute
cap r2_prof_ntt ntt_pass "-c 2" python tools/profile_run.py --reps 1 --ntt-cols 8 --what ntt
B2_NTT_VARIANT=0 cap r2_prof_ntt_v0 ntt_pass "-c 2" python tools/profile_run.py --reps 1 --ntt-cols 8 --what ntt
cap r2_prof_quot quotient_eval "-c 1" python tools/quotient_bench.py --k 20 --reps 1
ls -la $O | tail -12
python bench.py --steps 10 --warmup 3 > $O/r2_bench_b.json 2> $O/r2_bench_b.err
tail -c 400 $O/r2_bench_b.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'blocking', d['e2e']['blocking_call']['value'], 'parity', d['parity_check']['ok'])
print('ntt', d['ntt']['value'])
print('proof18', d['create_proof'].get('value'), 'proof22', d['create_proof_k22'].get('value'), d['create_proof_k22'].get('phases_s'), d['create_proof_k22'].get('error'))
PY
