#!/bin/bash
# ncu --set full of one MOMENT-FREE three-update launch on the C4 lattice as lb_step launches it by itself (128-row
# segments, 32-row ones at the end): tb2_sweep --shapes auto --steps 6 -> the first fused_march launch of the process
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 0 -c 1 -o $O/r2_final2_ncu_march3_f32_strict_c4 \
   python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 6 --reps 1 --shapes auto > $O/r2_final2_ncu3_c4.log 2>&1
f=$O/r2_final2_ncu_march3_f32_strict_c4; ncu -i $f.ncu-rep --page details > ${f}_details.txt 2>/dev/null; ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Issue Slots Busy|Executed Ipc|Grid Size" ${f}_details.txt | head
rm -f $f.ncu-rep
