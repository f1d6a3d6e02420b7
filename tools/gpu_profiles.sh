#!/bin/bash
# Reproduces the single-GPU evidence under profiles/ (run with: gpurun -- 'bash tools/gpu_profiles.sh').
#   1. tile-configuration sweeps (fp32/fp64, STRICT/FAST, register-shuffle and TMA variants)
#   2. ncu --set full of one moment-free launch per kernel flavour (DRAM bytes, instruction counts)
#   3. launch list of the bench command
#   4. the bench lines of every workload + the CPU reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in strict fast; do
  timeout 400 python tools/sweep.py --dtype f32 --math $m --bc pipe --mask > gpurun_out/sweep_f32_$m.txt 2>&1
  timeout 400 python tools/sweep.py --dtype f64 --math $m --bc pipe --mask --nx 16384 --ny 8192 > gpurun_out/sweep_f64_$m.txt 2>&1
done
for d in f32 f64; do for m in strict fast; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -o gpurun_out/prof_${d}_$m \
     python tools/sweep.py --dtype $d --math $m --bc pipe --mask --nx 8192 --ny 8192 --steps 2 --reps 1 \
     --filter $d.$m.v$([ $d = f32 ] && echo 4 || echo 2).wx2.wy2.r1.b6.ld1.st0 > gpurun_out/ncu_${d}_$m.log 2>&1
done; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c4.csv \
   python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4_n1.err
for w in c1 c2 c3 c5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps $([ $w = c1 ] && echo 2000 || echo 200) --warmup 20 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 100 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
grep -h BEST gpurun_out/sweep_*.txt
cut -c1-200 gpurun_out/bench_c4_n1.json
# throughput of the Cython-order kernels (parity vehicles), 8192x8192 fp32
python - <<'PY' > gpurun_out/cython_scheme_throughput.txt 2>&1
import sys, time
sys.path.insert(0, "2d-lb_b200")
import numpy as np, torch
from lb_b200 import Lattice
for scheme, bc in (("cython", "pipe"), ("cython_old", "pipe"), ("cython_old", "velocity_yperiodic"),
                   ("opencl_old", "velocity_yperiodic")):
    s = torch.cuda.Stream()
    sim = Lattice(8192, 8192, 1.2, 1.01, 1.0, scheme=scheme, bc=bc, u_west=0.05, u_east=0.05, stream=s.cuda_stream)
    w = np.array([4/9] + [1/9]*4 + [1/36]*4, dtype=np.float32)
    sim.upload_f(np.broadcast_to(w[:, None, None], (9, 8192, 8192)))
    sim.run(5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(); sim.run(40, sync=False); e1.record()
    sim.sync()
    ms = e0.elapsed_time(e1) / 40
    print(f"{scheme:10s} {bc:18s} {ms:.4f} ms/step {8192*8192/ms/1e3:9.0f} MLUPS {8192*8192*72/ms/1e6:7.0f} GB/s")
    sim.close()
PY
cat gpurun_out/cython_scheme_throughput.txt
