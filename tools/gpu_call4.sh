#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for m in fast strict; do
  timeout 400 python tools/sweep.py --dtype f32 --math $m --bc pipe --mask > gpurun_out/sweep_f32_$m.txt 2>&1
  timeout 400 python tools/sweep.py --dtype f64 --math $m --bc pipe --mask --nx 16384 --ny 8192 > gpurun_out/sweep_f64_$m.txt 2>&1
done
for m in fast strict; do for d in f32 f64; do
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fused_step -s 3 -c 1 --csv --log-file gpurun_out/inst_${d}_$m.csv \
   python tools/sweep.py --dtype $d --math $m --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > /dev/null 2>&1
done; done
tail -n 4 gpurun_out/pytest_gpu.txt
grep -h BEST gpurun_out/sweep_*.txt
grep -h "inst_executed\|issue_active\|duration" gpurun_out/inst_*.csv | cut -d, -f5,13,15 
