#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/march_debug.py 2>&1 | grep -v identical | cut -c1-300 > gpurun_out/r2_march_debug2.txt; echo "debug lines not identical: $(wc -l < gpurun_out/r2_march_debug2.txt)"; head -5 gpurun_out/r2_march_debug2.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or two_update or self_ring or halo_timeout or half_as_many or slab_decomposition or strict_pipe_bitexact" > gpurun_out/r2_second_tests.txt 2>&1
tail -15 gpurun_out/r2_second_tests.txt
S=march.w4b4.s64,march.w4b4.s128,march.w4b3.s64,march.w8b2.s64,march.w2b8.s64,march.w4b4.scalar.s64
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes $S > gpurun_out/r2_second_sweep_c4.txt 2>&1
cat gpurun_out/r2_second_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes march.w4b4.s64,march.w4b3.s64 > gpurun_out/r2_second_sweep_c3.txt 2>&1
cat gpurun_out/r2_second_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --math fast --steps 20 --reps 2 --shapes march.w4b4.s64,march.w4b3.s64,march.w4b4.scalar.s64 > gpurun_out/r2_second_sweep_fast.txt 2>&1
cat gpurun_out/r2_second_sweep_fast.txt
