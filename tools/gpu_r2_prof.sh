#!/bin/bash
# ncu --set full with source of the shipped two-update kernel (8192^2: quick) + its opcode/stall tables
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHAPE=${1:-march.w4b5.sh.s64}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 2 -c 1 -o gpurun_out/r2_ncu_${SHAPE}_f32_strict_8192 \
   python tools/tb2_sweep.py --nx 8192 --ny 8192 --steps 4 --reps 1 --shapes $SHAPE > gpurun_out/r2_ncu_8192.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
