#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/march_debug.py 2>&1 | grep -v identical | cut -c1-300 > gpurun_out/r2_k3_debug.txt; echo "debug lines not identical: $(wc -l < gpurun_out/r2_k3_debug.txt)"; head -8 gpurun_out/r2_k3_debug.txt
S=march.w4b6.sh.bf.s64,march3.w4b5.s64,march3.w4b5.s32,march3.w4b4.s64,march3.w4b6.s64,march3.w4b5.s16
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 24 --reps 3 --shapes $S 2>&1 | tee gpurun_out/r2_k3_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 24 --reps 2 --shapes march.w4b5.sh.s32,march3.w4b5.s32,march3.w4b5.s64,march3.w4b4.s32 2>&1 | tee gpurun_out/r2_k3_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 24 --reps 2 --shapes march.w4b6.sh.bf.s16,march3.w4b5.s16,march3.w4b5.s32 2>&1 | tee gpurun_out/r2_k3_sweep_slab.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or two_update or self_ring or half_as_many or streamed or slab or halo_timeout" > gpurun_out/r2_k3_tests.txt 2>&1
tail -12 gpurun_out/r2_k3_tests.txt
