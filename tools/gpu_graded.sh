#!/bin/bash
# Short segments at the end of automatic launches: the new test, then automatic (graded) vs the same shape by name (uniform)
# on C4, its N=8 slab shape, C5 and C3's grid.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_graded_segments.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "short_segments or few_waves or temporal_blocking_is or tall_lattice or streamed" > gpurun_out/r2_graded_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_graded_pytest.txt
tail -n 4 gpurun_out/r2_graded_pytest.txt
: > $O
for round in 1 2; do
timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 24 --reps 5 --shapes march3.w4b4.s64,auto 2>&1 | grep -v "^off" | cut -c1-130 >> $O
timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64,auto 2>&1 | grep -v "^off" | cut -c1-130 >> $O
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --reps 3 --shapes march3.w4b5.s64,auto 2>&1 | grep -v "^off" | cut -c1-130 >> $O
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 24 --reps 3 --shapes march3.w4b4.s64,auto 2>&1 | grep -v "^off" | cut -c1-130 >> $O
done
cat $O
