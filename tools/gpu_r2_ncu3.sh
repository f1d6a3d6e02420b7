#!/bin/bash
# ncu --set full of one MOMENT-FREE three-update launch on the C4 lattice (tb2_sweep --steps 6: march launches are
# [3, 3+moments] per run, -s 2 skips the first run's two)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 2 -c 1 -o $O/r2_final_ncu_march3_f32_strict_c4 \
   python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 6 --reps 1 --shapes auto > $O/r2_final_ncu3_c4.log 2>&1
f=$O/r2_final_ncu_march3_f32_strict_c4; ncu -i $f.ncu-rep --page details > ${f}_details.txt 2>/dev/null; ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Issue Slots Busy|Executed Ipc" ${f}_details.txt | head
