#!/bin/bash
# Builds A/B variants of libhmp_planner.so into tools/_ab/<name>/ (git-ignored, travels to the GPU box):
#   tools/ab_build.sh name "-DHMP_TPC_THREADS=448 -DHMP_TPC_MIN_BLOCKS=1" [name2 "flags2" ...]
# Run a variant with HMP_LIB=tools/_ab/<name>/libhmp_planner.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p tools/_ab/$name
  ( cd humap_local_planner_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
      -Xcompiler -fPIC -Xcompiler -pthread -shared $flags -Xptxas -v -o ../../tools/_ab/$name/libhmp_planner.so hmp_kernels.cu hmp_api.cu \
      2> ../../tools/_ab/$name/ptxas.log ) &
done
wait
for d in tools/_ab/*/; do echo "== $d"; grep -A1 "sweep_tpc_kernel" $d/ptxas.log | grep -E "registers|spill" | head -4; done
