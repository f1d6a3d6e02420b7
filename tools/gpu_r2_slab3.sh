#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dims in "4096 32768" "8192 32768" "16384 32768"; do
  set -- $dims
  timeout 300 python tools/tb2_sweep.py --nx $1 --ny $2 --steps 24 --reps 3 --shapes march3.w4b4.s16,march3.w4b4.s32,march3.w4b4.s64,march3.w4b5.s32,march3.w4b5.s64,march.w4b6.sh.bf.s16
done 2>&1 | tee gpurun_out/r2_slab3_sweep.txt
