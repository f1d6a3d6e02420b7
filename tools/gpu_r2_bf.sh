#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=march.w4b5.sh.s64,march.w4b5.sh.bf.s64,march.w4b6.sh.bf.s64,march.w4b4.sh.bf.s64
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 3 --shapes $S 2>&1 | tee gpurun_out/r2_bf_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_bf_sweep_c5.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_bf_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --math fast --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_bf_sweep_fast.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or two_update or bulky" > gpurun_out/r2_bf_tests.txt 2>&1
tail -5 gpurun_out/r2_bf_tests.txt
