#!/bin/bash
# Second GPU call: full GPU test-suite, error-growth table, bench lines, ncu traffic captures, launch list.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python tools/measure_errors.py > gpurun_out/errors.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err; echo "rc=$?" >> gpurun_out/bench_c4_n1.err
timeout 300 python bench.py --workload c2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload c3 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 300 python bench.py --workload c5 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 300 python bench.py --math strict --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_strict.json 2> gpurun_out/bench_c4_strict.err
timeout 600 python bench.py --impl reference --steps 100 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# ncu: moment-free fused launch (4th fused launch = first of the timed run), fp32 and fp64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_r1_f32_fast \
   python tools/sweep.py --dtype f32 --math fast --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_f32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_r1_f64_fast \
   python tools/sweep.py --dtype f64 --math fast --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_f64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_r1_f32_strict \
   python tools/sweep.py --dtype f32 --math strict --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_f32s.log 2>&1
# launch list of the bench command (every launch with its device time)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c4.csv \
   python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.txt
cat gpurun_out/bench_c4_n1.json
