#!/usr/bin/env python
"""L2-level temporal blocking experiment: lb_step_banded vs lb_step on one GPU.

    python tools/banded_sweep.py --nx 4096 --ny 32768 [--dtype f32 --steps 41]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2d-lb_b200"))
import numpy as np
import torch
from lb_b200 import Lattice


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=32768)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--steps", type=int, default=41)
    ap.add_argument("--bands", default="16,32,64,128,256")
    ap.add_argument("--depths", default="2,4,8")
    a = ap.parse_args()
    dtype = np.float32 if a.dtype == "f32" else np.float64
    s = torch.cuda.Stream()
    sim = Lattice(a.nx, a.ny, 1.7, 1.003, 1.0, dtype=dtype, stream=s.cuda_stream)
    sim.set_mask_disk(a.nx / 4.0, a.ny / 2.0, a.ny / 10.0)
    sim.init_synthetic("pipe_ramp", u0=0.05, amplitude=1e-3, seed=2015)
    row_mb = a.nx * 9 * (4 if a.dtype == "f32" else 8) / 1e6

    def timed(fn):
        fn()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                e0.record(); fn(); e1.record()
            sim.sync()
            best = min(best, e0.elapsed_time(e1) / a.steps)
        return best

    base = timed(lambda: sim.run(a.steps, sync=False))
    print(f"# {a.nx}x{a.ny} {a.dtype} strict, {a.steps} steps per run; one row of nine planes = {row_mb:.2f} MB")
    print(f"lb_step                      {base:8.4f} ms/step {a.nx * a.ny / base / 1e3:9.0f} MLUPS  x1.000", flush=True)
    for h in [int(v) for v in a.bands.split(",")]:
        for k in [int(v) for v in a.depths.split(",")]:
            ms = timed(lambda: sim.run_banded(a.steps, h, k, sync=False))
            print(f"banded rows {h:4d} ({h * row_mb:6.1f} MB) depth {k:2d}  {ms:8.4f} ms/step {a.nx * a.ny / ms / 1e3:9.0f} MLUPS  "
                  f"x{base / ms:5.3f}", flush=True)
    sim.close()


if __name__ == "__main__":
    main()
