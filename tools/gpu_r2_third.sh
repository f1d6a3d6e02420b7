#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_third_pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2_third_pytest_gpu.txt
for pf in 0 2 4 8; do
  echo "== LB_MARCH_PREFETCH=$pf"
  LB_MARCH_PREFETCH=$pf timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes march.w4b4.s64,march.w4b4.s128 2>&1 | grep march
done > gpurun_out/r2_third_prefetch.txt 2>&1
cat gpurun_out/r2_third_prefetch.txt
bash tools/gpu_r2_ncu.sh march.w4b4.s64
