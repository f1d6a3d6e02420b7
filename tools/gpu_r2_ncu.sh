#!/bin/bash
# ncu captures of the shipped two-update kernel: full set on 8192^2 (stall reasons, instruction counts) and on the
# C4 lattice (DRAM bytes per launch for roofline.traffic), plus the launch list of the bench command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHAPE=${1:-march.w4b4.s64}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 2 -c 1 -o gpurun_out/r2_ncu_march_f32_strict_8192 \
   python tools/tb2_sweep.py --nx 8192 --ny 8192 --steps 4 --reps 1 --shapes $SHAPE > gpurun_out/r2_ncu_march_8192.log 2>&1
timeout 1500 ncu --set full --clock-control none -k regex:fused_march -s 2 -c 1 -o gpurun_out/r2_ncu_march_f32_strict_c4 \
   python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 4 --reps 1 --shapes $SHAPE > gpurun_out/r2_ncu_march_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
