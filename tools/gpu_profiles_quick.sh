#!/bin/bash
# Refreshes the ncu evidence of the default kernel only (fp32 STRICT): one --set full capture of a moment-free
# launch at 8192^2 and the launch list of the bench command.  The full set: tools/gpu_profiles.sh.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_f32_strict \
   python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 \
   --filter f32.strict.v4.wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_f32_strict.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c4.csv \
   python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ls -la gpurun_out/prof_f32_strict.ncu-rep gpurun_out/launches_bench_c4.csv
grep -c fused_step gpurun_out/launches_bench_c4.csv
