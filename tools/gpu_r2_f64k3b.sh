#!/bin/bash
# 128-row segments for three-update launches (fp32 C4, fp64 C5), ncu --set full of one moment-free fp64 three-update
# launch on the C5 lattice, the C5 and C4 bench lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64,march3.w4b4.s128,march3.w4b5.s128,auto > $O/r2_march3_s128_c4.txt 2>&1
cat $O/r2_march3_s128_c4.txt
timeout 600 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 36 --reps 3 --shapes march3.w4b5.s64,march3.w4b5.s128,march3.w4b4.s128,auto > $O/r2_march3_s128_c5.txt 2>&1
cat $O/r2_march3_s128_c5.txt
timeout 600 python tools/tb2_sweep.py --nx 4096 --ny 32768 --steps 24 --reps 3 --shapes march3.w4b4.s64,march3.w4b4.s128,march3.w4b5.s64,auto > $O/r2_march3_s128_slab.txt 2>&1
cat $O/r2_march3_s128_slab.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fused_march -s 2 -c 1 -o $O/r2_final_ncu_march3_f64_strict_c5 \
   python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 6 --reps 1 --shapes march3.w4b5.s64 > $O/r2_final_ncu3_c5.log 2>&1
f=$O/r2_final_ncu_march3_f64_strict_c5; ncu -i $f.ncu-rep --page details > ${f}_details.txt 2>/dev/null; ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null
grep -E "Duration|DRAM Throughput|Registers Per|Issue Slots Busy|Executed Ipc|dram__bytes_(read|write).sum " ${f}_details.txt ${f}_raw.csv | head
timeout 600 python bench.py --workload c5 --steps 30 --warmup 5 --no-cpu-baseline > $O/r2_f64k3_bench_c5.json 2> $O/r2_f64k3_bench_c5.err; tail -c 1500 $O/r2_f64k3_bench_c5.json
