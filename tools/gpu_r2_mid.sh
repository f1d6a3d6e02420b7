#!/bin/bash
# mid-size lattices: which segment height (if any) beats the graph-batched one-update kernel?
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B=march.w4b6.sh.bf
for dims in "4096 1024" "3751 1251" "8192 2048" "8192 4096" "16384 4096" "16384 8192"; do
  set -- $dims
  timeout 300 python tools/tb2_sweep.py --nx $1 --ny $2 --steps 40 --reps 3 --shapes $B.s8,$B.s16,$B.s32,$B.s64,$B.s128
done 2>&1 | tee gpurun_out/r2_mid_sweep.txt
timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes $B.s64,$B.s128,march.w4b5.sh.bf.s64,march.w4b5.sh.bf.s128 2>&1 | tee gpurun_out/r2_mid_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 8192 --ny 2048 --no-mask --steps 40 --reps 3 --shapes march.w4b5.sh.s8,march.w4b5.sh.s16,march.w4b5.sh.s32,march.w4b5.sh.s64 2>&1 | tee gpurun_out/r2_mid_sweep_nomask.txt
timeout 300 python tools/tb2_sweep.py --nx 8192 --ny 4096 --dtype f64 --no-mask --steps 40 --reps 3 --shapes march.w4b5.sh.s8,march.w4b5.sh.s16,march.w4b5.sh.s32,march.w4b5.sh.s64 2>&1 | tee -a gpurun_out/r2_mid_sweep_nomask.txt
