#!/bin/bash
# First GPU call: box facts, smoke, parity tests, variant sweeps, one ncu capture.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== box"; nvidia-smi -L; nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null; python -c "import os; print('cpus', os.cpu_count())"
  lscpu | grep -E "Model name|Socket|Thread|Core" 
} > gpurun_out/box.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.txt
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 600 python tools/sweep.py --dtype f32 --math fast --bc pipe --mask --out gpurun_out/sweep_f32_fast.json > gpurun_out/sweep_f32_fast.txt 2>&1
timeout 300 python tools/sweep.py --dtype f32 --math strict --bc pipe --mask > gpurun_out/sweep_f32_strict.txt 2>&1
timeout 600 python tools/sweep.py --dtype f64 --math fast --bc pipe --mask --nx 16384 --ny 8192 > gpurun_out/sweep_f64_fast.txt 2>&1
timeout 300 python tools/sweep.py --dtype f64 --math strict --bc pipe --mask --nx 16384 --ny 8192 > gpurun_out/sweep_f64_strict.txt 2>&1
timeout 300 python tools/sweep.py --dtype f32 --math fast --bc periodic --filter wx2.wy2.r1 > gpurun_out/sweep_f32_periodic.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 4 -c 2 -o gpurun_out/prof_r1_default \
   python tools/sweep.py --dtype f32 --math fast --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 --filter wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_log.txt 2>&1
tail -3 gpurun_out/smoke.txt gpurun_out/pytest_gpu.txt
grep BEST gpurun_out/sweep_*.txt
