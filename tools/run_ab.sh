#!/bin/bash
# A/B of an environment switch on the GPU box: tools/run_ab.sh <tag> <VAR=value> -> two bench lines (default, switch set)
set -u
OUT=gpurun_out; TAG=$1; SW=$2
mkdir -p $OUT
B="python bench.py --steps 10 --seeds 3 --no-config4 --no-exact --no-cpu-baseline"
$B > $OUT/bench_abA_$TAG.json 2> $OUT/bench_ab_$TAG.err
env $SW $B > $OUT/bench_abB_$TAG.json 2>> $OUT/bench_ab_$TAG.err
for f in $OUT/bench_abA_$TAG.json $OUT/bench_abB_$TAG.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('cycle', round(d['ms_per_step'],3), 'sweep', [round(r['sweep_ms'],3) for r in d['per_seed']], 'sel', d['selection_matches_reference'])" 2>&1 | tail -1)"; done
