#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --nx 4096 --ny 32768 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/bench_shape_4096x32768.json 2> gpurun_out/bench_shape.err
timeout 300 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 4096 --ny 32768 --steps 50 > gpurun_out/sweep_f32_strict_4096x32768.txt 2>&1
tail -n 15 gpurun_out/pytest_gpu.txt | cut -c1-300
cut -c1-300 gpurun_out/bench_shape_4096x32768.json
cat gpurun_out/sweep_f32_strict_4096x32768.txt
