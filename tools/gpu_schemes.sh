#!/bin/bash
# Throughput of the parity-vehicle kernels (Cython-order and OLD-OpenCL-order schemes), 8192x8192 fp32.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/scheme_throughput.txt 2>&1
import sys
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
cat gpurun_out/scheme_throughput.txt
