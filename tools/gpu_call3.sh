#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask > gpurun_out/sweep_f32_strict.txt 2>&1
timeout 300 python tools/sweep.py --dtype f64 --math strict --bc pipe --mask --nx 16384 --ny 8192 > gpurun_out/sweep_f64_strict.txt 2>&1
timeout 300 python bench.py --math strict --steps 50 --no-cpu-baseline --no-e2e --variant f32.strict.v4.wx2.wy2.r1.b6.ld1.st0 > gpurun_out/bench_c4_strict.json 2> gpurun_out/bench_c4_strict.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_r1_f32_strict2 \
   python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_f32s.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.txt
cat gpurun_out/sweep_f32_strict.txt gpurun_out/sweep_f64_strict.txt gpurun_out/bench_c4_strict.json
