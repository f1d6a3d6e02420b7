#!/bin/bash
# Both orders of measurement (automatic choice first / last in the process): the second measurement of a process runs 4 % slower
# at the power cap, whichever kernel it is.
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; O=gpurun_out/r2_graded_segments_order.txt; : > $O
for r in 1 2; do
python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 24 --reps 3 --shapes march3.w4b4.s64,auto --auto-first 2>&1 | grep -v "^off" | cut -c1-140 >> $O
python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 24 --reps 3 --shapes march3.w4b4.s64,auto 2>&1 | grep -v "^off" | cut -c1-140 >> $O
done
python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 24 --reps 5 --shapes march3.w4b4.s64,auto --auto-first 2>&1 | grep -v "^off" | cut -c1-140 >> $O
python tools/tb2_sweep.py --nx 8192 --ny 32768 --steps 24 --reps 5 --shapes march3.w4b4.s64,auto --auto-first 2>&1 | grep -v "^off" | cut -c1-140 >> $O
python tools/tb2_sweep.py --nx 16384 --ny 32768 --steps 24 --reps 5 --shapes march3.w4b4.s64,auto --auto-first 2>&1 | grep -v "^off" | cut -c1-140 >> $O
cat $O
