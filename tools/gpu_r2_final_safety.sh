#!/bin/bash
# Final tree: compute-sanitizer memcheck over every compiled marching shape on small lattices (tools/tb2_sanitize.py, fp32
# and fp64 incl. the fp64 three-update shapes) and over the halo tests (two-/three-update slabs, narrow last strips,
# self-ring, timeout); then the ncu launch list of the driver's bench command.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/tb2_sanitize.py > $O/r2_final2_memcheck_shapes.txt 2>&1; echo "rc=$?" >> $O/r2_final2_memcheck_shapes.txt
tail -n 4 $O/r2_final2_memcheck_shapes.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "two_update or self_ring or half_as_many or halo_timeout or narrower_than" > $O/r2_final2_memcheck_halo.txt 2>&1; echo "rc=$?" >> $O/r2_final2_memcheck_halo.txt
tail -n 5 $O/r2_final2_memcheck_halo.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final2_launches_bench_c4.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2_final2_launches_bench_c4.log 2>&1
grep -c "fused_march" $O/r2_final2_launches_bench_c4.csv
