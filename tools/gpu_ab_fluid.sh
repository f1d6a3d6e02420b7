#!/bin/bash
# A/B: work items without a solid node run the obstacle-free copy of the march (working tree) vs the previous commit's
# library (build/liblb_d2q9_base.so, built from `git archive HEAD` where nvcc is).  Parity tests first.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_ab_fluid_only.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking_is or halo or slab or short_segments or few_waves or streamed or obstacle" > gpurun_out/r2_ab_fluid_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_ab_fluid_pytest.txt
tail -n 4 gpurun_out/r2_ab_fluid_pytest.txt | cut -c1-300
: > $O
for round in 1 2; do
  for lib in base new; do
    if [ $lib = base ]; then export LB_D2Q9_LIB=$PWD/build/liblb_d2q9_base.so; else unset LB_D2Q9_LIB; fi
    echo "## $lib, round $round" >> $O
    timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64 2>&1 | grep "^march" | cut -c1-160 >> $O
    timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --reps 3 --shapes march3.w4b5.s64 2>&1 | grep "^march" | cut -c1-160 >> $O
    timeout 300 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64 2>&1 | grep "^march" | cut -c1-160 >> $O
  done
done
cat $O
