#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
run() {  # name, env...
  name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes march.w4b4.s64 2>&1 | grep march
  env "$@" timeout 600 ncu --metrics $M --clock-control none -k regex:fused_march -s 2 -c 1 --csv --log-file gpurun_out/r2_fifth_$name.csv \
      python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 4 --reps 1 --shapes march.w4b4.s64 > /dev/null 2>&1
  grep -E "dram__bytes|duration|hit_rate" gpurun_out/r2_fifth_$name.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
}
run evict_last A=1
run plain LB_D2Q9_LIB=$PWD/build/liblb_rimld0.so
