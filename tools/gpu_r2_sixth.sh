#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/march_debug.py 2>&1 | grep -v identical | cut -c1-300 > gpurun_out/r2_lap_debug.txt; echo "debug lines not identical: $(wc -l < gpurun_out/r2_lap_debug.txt)"; head -12 gpurun_out/r2_lap_debug.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "temporal_blocking or two_update or self_ring or halo_timeout or half_as_many" > gpurun_out/r2_sixth_tests.txt 2>&1
tail -12 gpurun_out/r2_sixth_tests.txt
S=march.w4b4.s64,lap.w4b4.s64,lap.w4b4.s128,lap.w8b2.s64,lap.w4b3.s64,lap.w4b4.scalar.s64
timeout 600 python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 20 --reps 2 --shapes $S 2>&1 | tee gpurun_out/r2_sixth_sweep_c4.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --bc periodic --steps 20 --reps 2 --shapes march.w4b4.s64,lap.w4b4.s64 2>&1 | tee gpurun_out/r2_sixth_sweep_c3.txt
timeout 300 python tools/tb2_sweep.py --nx 16384 --ny 16384 --dtype f64 --steps 20 --reps 2 --shapes march.w4b4.s64,lap.w4b4.s64,lap.w8b2.s64 2>&1 | tee gpurun_out/r2_sixth_sweep_c5.txt
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum"
for sh in lap.w4b4.s64 lap.w8b2.s64; do
timeout 600 ncu --metrics $M --clock-control none -k regex:fused_march -s 2 -c 1 --csv --log-file gpurun_out/r2_sixth_$sh.csv \
      python tools/tb2_sweep.py --nx 32768 --ny 32768 --steps 4 --reps 1 --shapes $sh > /dev/null 2>&1
echo $sh; grep -E "dram__bytes|duration|hit_rate|inst_exec" gpurun_out/r2_sixth_$sh.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
