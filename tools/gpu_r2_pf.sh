#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=march.w4b4.s64,march.w4b3.pf.s64,march.w4b4.pf.s64,march.w8b1.pf.s64,march.w4b3.pf.s128
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_pf_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 20 --reps 2 --shapes march.w4b4.s64,march.w4b3.pf.s64,march.w4b4.pf.s64 2>&1 | tee gpurun_out/r2_pf_sweep_c5.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes march.w4b4.s64,march.w4b3.pf.s64 2>&1 | tee gpurun_out/r2_pf_sweep_c3.txt
