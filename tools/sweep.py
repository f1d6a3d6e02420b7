#!/usr/bin/env python
"""Time every compiled tile configuration of the fused kernel on one GPU (tuning aid).

    python tools/sweep.py --nx 16384 --ny 16384 --dtype f32 --math fast --bc pipe --steps 20

Prints one line per variant: MLUPS, achieved GB/s (72 or 144 B per lattice update) and the
fraction of MEASURED_PEAKS.json's hbm_gbs.  Timing: CUDA events on the launch stream, 3 warm-up
steps, working set >> L2.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2d-lb_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from lb_b200 import Lattice, native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--math", default="fast")
    ap.add_argument("--bc", default="pipe")
    ap.add_argument("--mask", action="store_true")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--filter", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()

    peak = 6538.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dtype = np.float32 if a.dtype == "f32" else np.float64
    bpl = 72 if a.dtype == "f32" else 144
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    sim = Lattice(a.nx, a.ny, 1.7, 1.003, 1.0, bc=a.bc, dtype=dtype, math=a.math, stream=stream.cuda_stream)
    if a.mask:
        sim.set_mask_disk(a.nx / 4, a.ny / 2, a.ny / 10)
    sim.init_synthetic("pipe_ramp" if a.bc == "pipe" else "shear_layers", u0=0.05, amplitude=1e-3, seed=1)
    sim.sync()
    rows = []
    names = [n for n in native.variants() if n.startswith(f"{a.dtype}.{a.math}.") and a.filter in n and ".d2q9i." not in n]
    for name in names:
        try:
            sim.set_variant(name)
        except native.LBError as exc:          # e.g. TMA variants on a periodic box
            print(f"{name:42s} skipped: {exc}", flush=True)
            continue
        sim.run(3)
        best = None
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record()
                sim.run(a.steps, sync=False)
                e1.record()
            sim.sync()
            ms = e0.elapsed_time(e1) / a.steps
            best = ms if best is None else min(best, ms)
        mlups = a.nx * a.ny / (best * 1e-3) / 1e6
        gbs = mlups * 1e6 * bpl / 1e9
        rows.append((name, best, mlups, gbs, gbs / peak))
        print(f"{name:42s} {best:8.4f} ms/step {mlups:10.0f} MLUPS {gbs:8.0f} GB/s {gbs / peak:6.3f} of measured {peak:.0f}", flush=True)
    rows.sort(key=lambda r: r[1])
    print("BEST", rows[0][0], f"{rows[0][2]:.0f} MLUPS", f"{rows[0][4]:.3f}")
    if a.out:
        with open(a.out, "w") as fh:
            json.dump([dict(variant=r[0], ms_per_step=r[1], mlups=r[2], gbs=r[3], frac=r[4]) for r in rows], fh, indent=1)
    sim.close()


if __name__ == "__main__":
    main()
