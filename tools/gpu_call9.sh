#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "tma or variants" > gpurun_out/pytest_tma.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tma.txt
timeout 400 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --filter tma > gpurun_out/sweep_tma_f32_strict.txt 2>&1
timeout 400 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --filter wx2.wy2.r1.b6 > gpurun_out/sweep_ref_f32_strict.txt 2>&1
timeout 400 python tools/sweep.py --dtype f32 --math fast --bc pipe --mask --filter tma > gpurun_out/sweep_tma_f32_fast.txt 2>&1
timeout 400 python tools/sweep.py --dtype f64 --math strict --bc pipe --mask --nx 16384 --ny 8192 --filter tma > gpurun_out/sweep_tma_f64_strict.txt 2>&1
timeout 400 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 32768 --ny 32768 --steps 60 --filter tma.v4.ty4.b6 > gpurun_out/sweep_tma_c4.txt 2>&1
timeout 400 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 32768 --ny 32768 --steps 60 --filter wx2.wy2.r1.b6 > gpurun_out/sweep_ref_c4.txt 2>&1
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:fused_step -s 3 -c 1 --csv --log-file gpurun_out/inst_tma_f32_strict.csv \
   python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter tma.v4.ty4.b6 > /dev/null 2>&1
tail -n 5 gpurun_out/pytest_tma.txt | cut -c1-250
cat gpurun_out/sweep_tma_f32_strict.txt gpurun_out/sweep_ref_f32_strict.txt gpurun_out/sweep_tma_f32_fast.txt gpurun_out/sweep_tma_f64_strict.txt gpurun_out/sweep_tma_c4.txt gpurun_out/sweep_ref_c4.txt | grep -v BEST
grep -h "inst_executed\|issue_active\|duration" gpurun_out/inst_tma_f32_strict.csv | cut -d, -f13,15
