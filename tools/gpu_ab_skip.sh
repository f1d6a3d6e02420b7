#!/bin/bash
# A/B: the branch-free bounce-back selects always executed (shipped) vs skipped by a warp-uniform branch when no node of
# the warp's row is solid (side build -DLB_BF_SKIP -> build/liblb_d2q9_skip.so)
# Build the side library first (where nvcc is):
#   PYTHONPATH=2d-lb_b200 python -c "from lb_b200 import build; build.build_library(extra_flags=['-DLB_BF_SKIP'], out='$PWD/build/liblb_d2q9_skip.so', tag='.skip')"
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_ab_bf_skip.txt
: > $O
for round in 1 2; do
  for lib in shipped skip; do
    if [ $lib = skip ]; then export LB_D2Q9_LIB=$PWD/build/liblb_d2q9_skip.so; else unset LB_D2Q9_LIB; fi
    echo "## $lib, round $round" >> $O
    timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64,march.w4b6.sh.bf.s64 2>&1 | grep "^march" | cut -c1-130 >> $O
    timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --reps 3 --shapes march3.w4b5.s64 2>&1 | grep "^march" | cut -c1-130 >> $O
    timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 1024 --steps 240 --reps 5 --shapes march.w4b6.sh.bf.s8 2>&1 | grep "^march" | cut -c1-130 >> $O
  done
done
cat $O
