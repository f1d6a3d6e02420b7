#!/bin/bash
# Quick single-GPU verification: smoke, full GPU test-suite, default bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
if [ "${1:-}" = "bench" ]; then
  ( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
  timeout 300 python bench.py --workload c2 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
fi
tail -n 3 gpurun_out/smoke.txt; tail -n 6 gpurun_out/pytest_gpu.txt | cut -c1-250
[ -f gpurun_out/bench_default.json ] && cut -c1-250 gpurun_out/bench_default.json && cat gpurun_out/bench_default.time && cut -c1-200 gpurun_out/bench_c2.json
