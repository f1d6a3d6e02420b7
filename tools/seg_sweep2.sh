#!/bin/bash
# Tall + short segment heights by environment (side build -DLB_SEG_ROWS_ENV, see seg_sweep.sh): 128/32 against 64/16, and
# 20 against 16 warps per SM, on C4 and on its N=8 slab shape.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_seg_sweep2.txt
: > $O
export LB_D2Q9_LIB=$PWD/build/liblb_d2q9_seg.so
run() { local nx=$1 ny=$2 shape=$3; timeout 300 python tools/tb2_sweep.py --nx $nx --ny $ny --steps 24 --reps 3 --shapes $shape 2>&1 | grep "^$shape" | cut -c1-160; }
for round in 1 2; do
for cfg in "64 16" "128 32" "96 24" "64 0"; do set -- $cfg
  for shape in march3.w4b4.s64 march3.w4b5.s64; do
    echo -n "slab 4096x32768 S=$1/$2 " >> $O; LB_SEG_ROWS=$1 LB_SEG_ROWS2=$2 run 4096 32768 $shape >> $O
  done
  echo -n "C4 S=$1/$2 " >> $O; LB_SEG_ROWS=$1 LB_SEG_ROWS2=$2 run 32768 32768 march3.w4b4.s64 >> $O
done
done
cat $O
