#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A=march.w4b5.sh.bf; B=march.w4b6.sh.bf; C=march.w4b5.sh
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 3 --shapes $A.s16,$A.s32,$A.s64,$B.s16,$B.s32,$B.s64 2>&1 | tee gpurun_out/r2_seg_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 20 --reps 3 --shapes $A.s16,$A.s32,$A.s64,$B.s16,$B.s32,$B.s64 2>&1 | tee gpurun_out/r2_seg_sweep_slab.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 3 --shapes $C.s16,$C.s32,$C.s64 2>&1 | tee gpurun_out/r2_seg_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --no-mask --steps 20 --reps 3 --shapes $C.s16,$C.s32,$C.s64 2>&1 | tee gpurun_out/r2_seg_sweep_c5.txt
