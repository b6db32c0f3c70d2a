#!/bin/bash
# One measurement pass on the GPU box: gpurun --timeout 2400 -- 'bash tools/gpu_run.sh r02a [tests] [ab] [prof] [mode1]'
# Everything lands under gpurun_out/ (tools/make_profiles.py turns it into the tracked summaries under profiles/).
set -u
TAG=${1:-rXX}; shift
WHAT=" $* "
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader; nproc
if [[ "$WHAT" == *" tests "* ]]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider ) > $OUT/tests_$TAG.log 2>&1
  tail -15 $OUT/tests_$TAG.log
  grep -c GATE $OUT/tests_$TAG.log
fi
if [[ "$WHAT" == *" bench "* ]]; then
  ( time python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -20 $OUT/bench_$TAG.err; }
  tail -c 1500 $OUT/bench_$TAG.json; echo; tail -4 $OUT/bench_$TAG.err
fi
if [[ "$WHAT" == *" ab "* ]]; then
  python bench.py --steps 10 --seeds 2 --no-config4 --no-exact --no-cpu-baseline > $OUT/bench_ab_base_$TAG.json 2>> $OUT/ab_$TAG.err
  for d in tools/_ab/*/; do
    n=$(basename $d)
    HMP_LIB=$d/libhmp_planner.so python bench.py --steps 10 --seeds 2 --no-config4 --no-exact --no-cpu-baseline > $OUT/bench_ab_${n}_$TAG.json 2>> $OUT/ab_$TAG.err
  done
  for f in $OUT/bench_ab_*_$TAG.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('cycle', round(d['ms_per_step'],3), 'sweep', [round(r['sweep_ms'],3) for r in d['per_seed']], 'sel', d['selection_matches_reference'])" 2>&1 | tail -1)"; done
fi
if [[ "$WHAT" == *" ab1 "* ]]; then
  B1="python bench.py --precise 1 --steps 6 --seeds 2 --no-config4 --no-exact --no-cpu-baseline"
  $B1 > $OUT/bench_ab1_base_$TAG.json 2>> $OUT/ab1_$TAG.err
  for d in tools/_ab/*/; do
    n=$(basename $d)
    HMP_LIB=$d/libhmp_planner.so $B1 > $OUT/bench_ab1_${n}_$TAG.json 2>> $OUT/ab1_$TAG.err
    HMP_LIB=$d/libhmp_planner.so python bench.py --steps 10 --seeds 2 --no-config4 --no-exact --no-cpu-baseline > $OUT/bench_ab_${n}_$TAG.json 2>> $OUT/ab1_$TAG.err
  done
  for f in $OUT/bench_ab1_*_$TAG.json $OUT/bench_ab_*_$TAG.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('cycle', round(d['ms_per_step'],3), 'sweep', [round(r['sweep_ms'],3) for r in d['per_seed']], 'sel', d['selection_matches_reference'])" 2>&1 | tail -1)"; done
fi
if [[ "$WHAT" == *" mode1 "* ]]; then
  python bench.py --precise 1 --steps 6 --seeds 2 --no-config4 --no-exact --no-cpu-baseline > $OUT/bench_mode1_$TAG.json 2> $OUT/bench_mode1_$TAG.err
  tail -c 700 $OUT/bench_mode1_$TAG.json; echo
fi
if [[ "$WHAT" == *" prof "* ]]; then
  B="python bench.py --steps 2 --seeds 1 --warmup 3 --no-config4 --no-exact --no-cpu-baseline"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/ncu_launches_$TAG.log 2>&1
  grep -c sweep_tpc $OUT/${TAG}_launches.csv
  ncu --set full --clock-control none --import-source on -k regex:sweep_tpc -s 4 -c 1 -f -o $OUT/${TAG}_full $B > $OUT/ncu_full_$TAG.log 2>&1
  ls -la $OUT/${TAG}_full.ncu-rep
fi
if [[ "$WHAT" == *" hole "* ]]; then
  python tools/selection_hole_stats.py --plans 120 > $OUT/selection_hole_$TAG.json 2> $OUT/selection_hole_$TAG.err
  python -c "import json; print(json.load(open('$OUT/selection_hole_$TAG.json'))['summary'])" || tail -5 $OUT/selection_hole_$TAG.err
fi
if [[ "$WHAT" == *" replay "* ]]; then
  python bench.py --cfg cfg4 > $OUT/bench_replay_$TAG.json 2> $OUT/bench_replay_$TAG.err
  tail -c 900 $OUT/bench_replay_$TAG.json; echo
fi
if [[ "$WHAT" == *" prof1tpc "* ]]; then   # the exact mode's thread-per-candidate FP64 sweep (r02zz)
  B="python bench.py --precise 1 --steps 1 --seeds 1 --warmup 3 --no-config4 --no-exact --no-cpu-baseline"
  ncu --set full --clock-control none --import-source on -k regex:sweep_tpc -s 2 -c 1 -f -o $OUT/${TAG}_full_mode1 $B > $OUT/ncu_full_mode1_$TAG.log 2>&1
  ncu -i $OUT/${TAG}_full_mode1.ncu-rep --page raw --csv > $OUT/${TAG}_mode1_raw.csv 2>/dev/null
  python tools/ncu_summary.py $OUT/${TAG}_mode1_raw.csv > $OUT/${TAG}_mode1_ncu_summary.txt
  ls -la $OUT/${TAG}_full_mode1.ncu-rep
fi
if [[ "$WHAT" == *" cfg1 "* ]]; then
  python bench.py --cfg cfg1 --no-cpu-baseline > $OUT/bench_${TAG}_cfg1.json 2> $OUT/bench_${TAG}_cfg1.err
  tail -c 400 $OUT/bench_${TAG}_cfg1.json; echo
fi
if [[ "$WHAT" == *" prof1 "* ]]; then
  B="python bench.py --precise 1 --steps 1 --seeds 1 --warmup 3 --no-config4 --no-exact --no-cpu-baseline"
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:plan_kernel<\(bool\)0, double, \(bool\)0' -s 2 -c 1 -f -o $OUT/${TAG}_full_mode1 $B > $OUT/ncu_full_mode1_$TAG.log 2>&1
  ls -la $OUT/${TAG}_full_mode1.ncu-rep
fi
