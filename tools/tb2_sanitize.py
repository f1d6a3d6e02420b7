#!/usr/bin/env python
"""Small temporal-blocking runs for compute-sanitizer (memcheck / racecheck): every tile shape, pipe with
obstacles on every edge and a periodic box, fp32 and fp64, compared with the one-step kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "2d-lb_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import numpy as np
from lb_b200 import Lattice, native
from oracle import oracle as orc
from util import periodic_case, pipe_case

L = native.lib()
names = [L.lb_tb2_shape_name(k).decode() for k in range(1, L.lb_tb2_shape_count())]
bad = 0
for dtype in (np.float32, np.float64):
    for bc, (nx, ny) in (("pipe", (300, 45)), ("periodic", (256, 37))):
        if bc == "pipe":
            f0, m = pipe_case(orc, nx, ny, dtype, mask="touching", seed=3)
        else:
            f0, m = periodic_case(orc, nx, ny, dtype, amplitude=1e-3, seed=3), None
        with Lattice(nx, ny, 1.4, 1.01, 1.0, mask=m, f0=f0, bc=bc, dtype=dtype) as sim:
            sim.set_temporal_blocking("off")
            sim.run(7)
            want = sim.download("f")
        for name in names:
            with Lattice(nx, ny, 1.4, 1.01, 1.0, mask=m, f0=f0, bc=bc, dtype=dtype) as sim:
                try:
                    sim.set_temporal_blocking(name)
                except native.LBError:
                    continue
                sim.run(7)
                ok = np.array_equal(sim.download("f"), want)
                bad += not ok
                print(dtype.__name__, bc, name, "ok" if ok else "MISMATCH", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
