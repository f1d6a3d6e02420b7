#!/usr/bin/env python
"""A/B timing of two builds of the package on the same box: device-resident C4-style step.

    python tools/ab_bench.py PKG_DIR_A PKG_DIR_B [--nx 32768 --ny 32768 --steps 60 --rounds 3]

Each round times A then B (fresh process each, so the two libraries never share an address space);
prints ms/step and MLUPS per run.  PKG_DIR is a directory holding `lb_b200/` (e.g. 2d-lb_b200).
"""
import argparse
import json
import subprocess
import sys

CHILD = r'''
import sys, json
sys.path.insert(0, sys.argv[1])
nx, ny, steps, zero_vel = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
import numpy as np, torch
from lb_b200 import Lattice
s = torch.cuda.Stream()
sim = Lattice(nx, ny, 1.7, 1.003, 1.0, stream=s.cuda_stream, zero_obstacle_velocity=bool(zero_vel))
sim.set_mask_disk(nx / 4.0, ny / 2.0, ny / 10.0)
sim.init_synthetic("pipe_ramp", u0=0.05, amplitude=1e-3, seed=2015)
sim.run(5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record(); sim.run(steps, sync=False); e1.record()
sim.sync()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"ms": ms, "mlups": nx * ny / ms / 1e3, "mass": sim.total_mass()}))
sim.close()
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pkgs", nargs="+")
    ap.add_argument("--nx", type=int, default=32768)
    ap.add_argument("--ny", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--zero-vel", type=int, default=0)
    a = ap.parse_args()
    for r in range(a.rounds):
        for pkg in a.pkgs:
            out = subprocess.run([sys.executable, "-c", CHILD, pkg, str(a.nx), str(a.ny), str(a.steps), str(a.zero_vel)],
                                 capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
            print(f"round {r} {pkg:28s} {line}", flush=True)


if __name__ == "__main__":
    main()
