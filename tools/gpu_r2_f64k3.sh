#!/bin/bash
# fp64 with three updates per launch (two overlap lanes per side) and the narrow-last-strip hand-shake:
# parity tests, then the sweep of the C5 lattice against the two-update shapes.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or halo or slab or launches or streamed" > gpurun_out/r2_f64k3_pytest.txt 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_f64k3_pytest.txt
tail -n 5 gpurun_out/r2_f64k3_pytest.txt
timeout 600 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --shapes off,auto,march.w4b5.sh.s32,march.w4b5.sh.s64,march3.w4b4.s16,march3.w4b4.s32,march3.w4b4.s64,march3.w4b5.s32,march3.w4b5.s64,march3.w4b6.s32,march3.w4b6.s64 > gpurun_out/r2_f64k3_sweep_c5.txt 2>&1
cat gpurun_out/r2_f64k3_sweep_c5.txt
timeout 600 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --no-mask --shapes off,march.w4b5.sh.s64,march3.w4b4.s32,march3.w4b4.s64,march3.w4b5.s64,march3.w4b6.s64 > gpurun_out/r2_f64k3_sweep_c5_nomask.txt 2>&1
cat gpurun_out/r2_f64k3_sweep_c5_nomask.txt
