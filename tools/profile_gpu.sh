#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/profile_gpu.sh r01e'): plain bench first, then the ncu launch list of the same
# command, then ONE --set full capture of the dominant kernel; CSV exports land in gpurun_out/ (copy summaries to profiles/).
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; exit 1; }
tail -c 600 $OUT/bench_$TAG.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
grep -c plan_kernel $OUT/launches_$TAG.csv
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 6 -c 1 -f -o $OUT/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/src_$TAG.csv 2>/dev/null
ls -la $OUT/prof_$TAG.ncu-rep $OUT/raw_$TAG.csv $OUT/src_$TAG.csv
