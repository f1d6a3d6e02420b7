#!/bin/bash
# Temporal-blocking evidence: sanitizers on small runs, tile sweeps (fp64, 32768^2 fp32, periodic), ncu DRAM bytes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/tb2_sanitize.py > gpurun_out/tb2_memcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/tb2_memcheck.txt
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/tb2_sanitize.py > gpurun_out/tb2_racecheck.txt 2>&1; echo "rc=$?" >> gpurun_out/tb2_racecheck.txt
tail -n 3 gpurun_out/tb2_memcheck.txt; tail -n 3 gpurun_out/tb2_racecheck.txt
timeout 200 python tools/tb2_sweep.py --dtype f64 --nx 16384 --ny 8192 --shapes rows6.w8,rows14.w8,rows6.w4,rows14.w4,128x8.t256 --reps 2 2>&1 | tee gpurun_out/tb2_sweep_f64_strict.txt
timeout 200 python tools/tb2_sweep.py --nx 32768 --ny 32768 --shapes rows6.w8,rows14.w8 --reps 2 --steps 21 2>&1 | tee gpurun_out/tb2_sweep_f32_strict_32768.txt
timeout 200 python tools/tb2_sweep.py --bc periodic --shapes rows6.w8,rows14.w8 --reps 2 2>&1 | tee gpurun_out/tb2_sweep_f32_periodic.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_two_step -s 2 -c 1 -o gpurun_out/prof_tb2_f32_strict \
   python tools/tb2_sweep.py --nx 8192 --ny 8192 --shapes rows14.w8 --reps 1 --steps 9 > gpurun_out/ncu_tb2.log 2>&1
ls -la gpurun_out/prof_tb2_f32_strict.ncu-rep
