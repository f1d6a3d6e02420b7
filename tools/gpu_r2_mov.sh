#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/collide_ceiling > gpurun_out/r2_collide_ceiling.txt 2>&1; head -6 gpurun_out/r2_collide_ceiling.txt
S=march.w4b5.sh.s64,march.w4b4.s64
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 3 --shapes $S 2>&1 | tee gpurun_out/r2_mov_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_mov_sweep_c5.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_mov_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 20 --reps 3 --shapes $S 2>&1 | tee gpurun_out/r2_mov_sweep_slab.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or two_update or self_ring or half_as_many or streamed or tall_lattice or fast_math_error or device_handle or bulky" > gpurun_out/r2_mov_tests.txt 2>&1
tail -5 gpurun_out/r2_mov_tests.txt
