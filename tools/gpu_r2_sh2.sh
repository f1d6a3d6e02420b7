#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=march.w4b4.s64,march.w4b4.sh.s64,march.w4b5.sh.s64
for lib in "" "$PWD/build/liblb_bcool.so"; do
echo "=== LB_D2Q9_LIB=$lib"
LB_D2Q9_LIB=$lib timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes $S 2>&1 | grep -v "^off"
LB_D2Q9_LIB=$lib timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 20 --reps 2 --shapes $S 2>&1 | grep -v "^off"
LB_D2Q9_LIB=$lib timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes $S 2>&1 | grep -v "^off"
done 2>&1 | tee gpurun_out/r2_sh2.txt
