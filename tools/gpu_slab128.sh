#!/bin/bash
# the N=4 and N=2 slab shapes of C4 under the 128-row rule: automatic choice vs uniform 64-row segments, one process each
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; O=gpurun_out/r2_slab_shapes_128.txt; : > $O
for dims in "8192 32768" "16384 32768"; do set -- $dims
  for sh in auto march3.w4b4.s64; do
    timeout 200 python tools/tb2_sweep.py --nx $1 --ny $2 --steps 24 --reps 3 --shapes $sh 2>&1 | grep -v "^off" | cut -c1-160 >> $O
  done
done
cat $O
