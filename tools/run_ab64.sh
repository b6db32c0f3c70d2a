#!/bin/bash
# A/B of the FP64 thread-per-candidate sweep (exact mode): seeds 0 and 6, the in-tree library and every variant under tools/_ab
run() { # name lib threads
  for s in 0 6; do
    echo -n "$1 threads=$3 seed=$s: "
    HMP_LIB=$2 HMP_F64_TPC_THREADS=$3 timeout 300 python tools/time_seed.py --seed $s --precise 1 --reps 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['cycle_sweep_ms'], d['best_index'], d['sweep_mode'])"
  done
}
run base humap_local_planner_b200/lib/libhmp_planner.so 256
for d in tools/_ab/*/; do run $(basename $d) $d/libhmp_planner.so 256; done
