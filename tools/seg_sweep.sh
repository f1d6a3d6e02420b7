#!/bin/bash
# Segment heights that are not powers of two (side build with -DLB_SEG_ROWS_ENV: build/liblb_d2q9_seg.so reads
# LB_SEG_ROWS): does a height that fills the last wave of CTAs pay?  The N=8 slab shape of C4 (three updates per
# launch), C2 (two), C4 with FAST math for the record.
# Build the side library first (where nvcc is):
#   PYTHONPATH=2d-lb_b200 python -c "from lb_b200 import build; build.build_library(extra_flags=['-DLB_SEG_ROWS_ENV'], out='$PWD/build/liblb_d2q9_seg.so', tag='.seg')"
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r2_seg_sweep.txt
: > $O
export LB_D2Q9_LIB=$PWD/build/liblb_d2q9_seg.so
run() {  # nx ny steps shape extra... ; env LB_SEG_ROWS set by caller
  local nx=$1 ny=$2 steps=$3 shape=$4; shift 4
  timeout 300 python tools/tb2_sweep.py --nx $nx --ny $ny --steps $steps --reps 5 --shapes $shape "$@" 2>&1 | grep "^$shape" | cut -c1-100
}
echo "# 4096x32768 f32 strict, march3.w4b4 (16 warps per SM)" >> $O
for s in 48 52 56 58 60 61 62 64 70 72; do echo -n "S=$s " >> $O; LB_SEG_ROWS=$s run 4096 32768 24 march3.w4b4.s64 >> $O; done
echo "# 4096x32768 f32 strict, march3.w4b5 (20 warps per SM)" >> $O
for s in 44 48 50 52 55 64 74 76; do echo -n "S=$s " >> $O; LB_SEG_ROWS=$s run 4096 32768 24 march3.w4b5.s64 >> $O; done
echo "# 4096x1024 f32 strict (C2), march.w4b6.sh.bf (24 warps per SM)" >> $O
for s in 6 7 8 10 11 12 13 14 16; do echo -n "S=$s " >> $O; LB_SEG_ROWS=$s run 4096 1024 240 march.w4b6.sh.bf.s8 >> $O; done
echo "# 4096x1024 f32 strict (C2), march.w4b5.sh (20 warps per SM)" >> $O
for s in 7 8 10 12 13 14 16; do echo -n "S=$s " >> $O; LB_SEG_ROWS=$s run 4096 1024 240 march.w4b5.sh.s8 >> $O; done
echo "# 4096x1024 f32 strict (C2), march3.w4b4" >> $O
for s in 8 12 16 18 20 22; do echo -n "S=$s " >> $O; LB_SEG_ROWS=$s run 4096 1024 240 march3.w4b4.s16 >> $O; done
cat $O
unset LB_D2Q9_LIB
timeout 300 python bench.py --steps 20 --warmup 5 --math fast --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_c4_fast.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c4_fast.json').read().strip().splitlines()[-1]); print("C4 FAST", round(d["value"]), d["config"]["kernel"][:50])
PY
