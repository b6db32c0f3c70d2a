#!/bin/bash
# A/B of the exact mode (precision 1) on the GPU box: the in-tree library against every variant under tools/_ab/
set -u
OUT=gpurun_out; TAG=$1
mkdir -p $OUT
B1="python bench.py --precise 1 --steps 6 --seeds 3 --no-config4 --no-cpu-baseline"
$B1 > $OUT/bench_ab1_base_$TAG.json 2>> $OUT/ab1_$TAG.err
for d in tools/_ab/*/; do
  n=$(basename $d)
  HMP_LIB=$d/libhmp_planner.so $B1 > $OUT/bench_ab1_${n}_$TAG.json 2>> $OUT/ab1_$TAG.err
done
for f in $OUT/bench_ab1_*_$TAG.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('cycle', round(d['ms_per_step'],3), 'sweep', [round(r['sweep_ms'],3) for r in d['per_seed']], 'sel', d['selection_matches_reference'])" 2>&1 | tail -1)"; done
