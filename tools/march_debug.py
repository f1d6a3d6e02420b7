#!/usr/bin/env python
"""Debug aid: where does a two-update shape differ from the one-update kernel?  Prints differing (population, y, x)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2d-lb_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lb_b200 import Lattice, native
from oracle import oracle as orc
from util import pipe_case, periodic_case

def run(bc, nx, ny, math, shape, mask, steps, dtype=np.float32, zv=False):
    if bc == "pipe":
        f0, m = pipe_case(orc, nx, ny, dtype, mask=mask, seed=nx)
    else:
        f0, m = periodic_case(orc, nx, ny, dtype, amplitude=1e-3, seed=ny), None
    kw = dict(mask=m, f0=f0, bc=bc, dtype=dtype, math=math, zero_obstacle_velocity=zv)
    with Lattice(nx, ny, 1.4, 1.01, 1.0, **kw) as a:
        a.run(steps); want = a.fields()
    with Lattice(nx, ny, 1.4, 1.01, 1.0, **kw) as b:
        try:
            b.set_temporal_blocking(shape)
        except native.LBError as e:
            print(bc, nx, ny, math, shape, "refused:", e); return
        b.run(steps); got = b.fields()
    for k in ("f", "rho", "u", "v"):
        d = np.argwhere(got[k] != want[k])
        tag = f"{bc} {nx}x{ny} {math} {shape} mask={mask} steps={steps} {np.dtype(dtype).name} [{k}]"
        if len(d) == 0:
            print(tag, "identical")
        else:
            print(tag, f"{len(d)} differ; first:", d[:12].tolist(), " xs:", sorted(set(d[:, -1].tolist()))[:20], " ys:", sorted(set(d[:, -2].tolist()))[:20],
                  " max|d|:", float(np.abs(got[k].astype(np.float64) - want[k]).max()))

for shape in ("march3.w4b5.s32", "march3.w4b4.s8"):
    run("pipe", 5, 4, "strict", shape, "none", 3)
    run("pipe", 300, 70, "strict", shape, "none", 3)
    run("pipe", 300, 70, "strict", shape, "touching", 3)
    run("pipe", 300, 70, "fast", shape, "touching", 9)
    run("pipe", 700, 41, "strict", shape, "bulky", 6, zv=True)
    run("periodic", 128, 5, "strict", shape, None, 3)
    run("periodic", 256, 37, "strict", shape, None, 6)
    run("periodic", 384, 70, "fast", shape, None, 8)
