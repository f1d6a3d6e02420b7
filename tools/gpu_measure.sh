#!/bin/bash
# Single-GPU check after a kernel change: GPU test-suite, instruction counts of the default kernels, C4 bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -n 3 gpurun_out/pytest_gpu.txt | cut -c1-200
for m in strict fast; do
  timeout 300 python tools/sweep.py --dtype f32 --math $m --bc pipe --mask --filter f32.$m.v4.wx2.wy2.r1.b6.ld1.st0 | head -1
  timeout 300 python tools/sweep.py --dtype f64 --math $m --bc pipe --mask --nx 16384 --ny 8192 --filter f64.$m.v2.wx2.wy2.r1.b6.ld1.st0 | head -1
  timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fused_step -s 3 -c 1 --csv --log-file gpurun_out/inst_f32_$m.csv \
     python tools/sweep.py --dtype f32 --math $m --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter f32.$m.v4.wx2.wy2.r1.b6.ld1.st0 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/inst_f32_$m.csv") if l.startswith('"'))]
d={r['Metric Name']:float(r['Metric Value'].replace(',','')) for r in rows}
print("f32 $m: %.1f instr per lattice update, issue active %.1f %%" % (d['smsp__inst_executed.sum']*32/8192/8192, d['smsp__issue_active.avg.pct_of_peak_sustained_active']))
PY
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline | cut -c1-170
