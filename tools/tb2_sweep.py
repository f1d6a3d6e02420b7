#!/usr/bin/env python
"""Throughput of the two-update kernel (csrc/lb_march.cuh) per compiled shape, against the one-update kernel,
with an exact check at full size: every shape restarts from the same device-side initial state and must end
with the one-update kernel's checksum (64-bit sum of the populations' bit patterns).

    python tools/tb2_sweep.py [--nx 16384 --ny 16384 --dtype f32 --math strict --steps 40 --bc pipe --shapes a,b]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2d-lb_b200"))
import numpy as np
import torch
from lb_b200 import Lattice, native


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--math", default="strict")
    ap.add_argument("--bc", default="pipe")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--no-mask", action="store_true")
    ap.add_argument("--shapes", default="", help="comma-separated tile names (default: all)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--auto-first", action="store_true", help="time the automatic choice before the named shapes")
    a = ap.parse_args()
    dtype = np.float32 if a.dtype == "f32" else np.float64
    elem = 4 if a.dtype == "f32" else 8
    L = native.lib()
    names = [L.lb_tb2_shape_name(k).decode() for k in range(L.lb_tb2_shape_count())]
    s = torch.cuda.Stream()
    sim = Lattice(a.nx, a.ny, 1.7, 1.003, 1.0, bc=a.bc, dtype=dtype, math=a.math, stream=s.cuda_stream)
    if a.bc == "pipe" and not a.no_mask:
        sim.set_mask_disk(a.nx / 4.0, a.ny / 2.0, a.ny / 10.0)
    print(f"# {a.nx}x{a.ny} {a.dtype} {a.math} {a.bc}, {a.steps} steps per run")
    base, want = None, None
    wanted = [w for w in a.shapes.split(",") if w]
    runs = [(k, name) for k, name in enumerate(names) if not wanted or name == "off" or name in wanted]
    if "auto" in wanted:                # what lb_step does on this lattice by itself (shape, segment heights)
        sim.set_temporal_blocking("auto")
        runs.insert(1 if a.auto_first else len(runs), (-1, f"auto:{sim.temporal_blocking.replace('march', 'm')}/{sim.segment_rows}"))
    for k, name in runs:
        try:
            sim.set_temporal_blocking(k)
        except native.LBError as exc:
            print(f"{name:22s} not available: {exc}")
            continue
        sim.init_synthetic("pipe_ramp" if a.bc == "pipe" else "shear_layers", u0=0.05, amplitude=1e-3, seed=2015)
        sim.run(a.steps + 1)            # odd: one single-update launch + two-update launches
        csum = sim.checksum()
        want = csum if want is None else want
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                e0.record(); sim.run(a.steps, sync=False); e1.record()
            sim.sync()
            best = min(best, e0.elapsed_time(e1) / a.steps)
        mlups = a.nx * a.ny / best / 1e3
        base = base or mlups
        print(f"{name:22s} {best:8.4f} ms/step {mlups:9.0f} MLUPS  x{mlups / base:5.3f}  "
              f"({mlups * 1e6 * 18 * elem / 1e9:7.0f} GB/s-equivalent at {18 * elem} B/LU)   "
              f"bits == one-update kernel: {'yes' if csum == want else 'NO'}", flush=True)
    sim.close()


if __name__ == "__main__":
    main()
