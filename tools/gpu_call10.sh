#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
for m in strict fast; do
timeout 300 python tools/sweep.py --dtype f32 --math $m --bc pipe --mask --filter wx2.wy2.r1.b6.ld1.st0 | head -1
timeout 300 python tools/sweep.py --dtype f64 --math $m --bc pipe --mask --nx 16384 --ny 8192 --filter wx2.wy2.r1.b6.ld1.st0 | head -1
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fused_step -s 3 -c 1 --csv --log-file gpurun_out/inst2_f32_$m.csv \
   python tools/sweep.py --dtype f32 --math $m --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > /dev/null 2>&1
grep -h "inst_executed\|issue_active" gpurun_out/inst2_f32_$m.csv | cut -d, -f13,15
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline | cut -c1-160
timeout 100 python bench.py --workload c1 --steps 2000 --warmup 100 --no-cpu-baseline | cut -c1-200
tail -n 3 gpurun_out/pytest_gpu.txt | cut -c1-200
