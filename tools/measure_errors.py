#!/usr/bin/env python
"""Prints FAST-vs-oracle error growth with step count (used to state tolerances in DESIGN.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-lb_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
from lb_b200 import Lattice  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from util import periodic_case, pipe_case  # noqa: E402

for label, dtype, bc, kw in (("pipe256x128 f32", np.float32, "pipe", dict(mask="blocks")),
                             ("shear256x128 u0=0.1 f32", np.float32, "periodic", {}),
                             ("pipe256x128 f64", np.float64, "pipe", dict(mask="blocks"))):
    if bc == "pipe":
        f0, m = pipe_case(orc, 256, 128, dtype, **kw)
        omega = 1.3
    else:
        f0, m = periodic_case(orc, 256, 128, dtype, u0=0.1), None
        omega = 1.5
    ref = orc.OpenCLSchemeOracle(f0, omega, 1.01, 1.0, mask=m, dtype=dtype,
                                 bc=orc.BC_PERIODIC if bc == "periodic" else orc.BC_PIPE)
    sim = Lattice(256, 128, omega, 1.01, 1.0, mask=m, f0=f0, dtype=dtype, math="fast", bc=bc)
    done = 0
    for n in (1, 10, 100, 500, 1000, 2000):
        ref.run(n - done)
        sim.run(n - done)
        done = n
        r, u = sim.download("rho"), sim.download("u")
        print(f"{label:26s} N={n:5d} max|drho|/max rho={np.abs(r - ref.rho).max() / np.abs(ref.rho).max():.3e} "
              f"max|du|={np.abs(u - ref.u).max():.3e} max|u|={np.abs(ref.u).max():.3e} "
              f"rel u={np.abs(u - ref.u).max() / np.abs(ref.u).max():.3e}", flush=True)
    sim.close()
