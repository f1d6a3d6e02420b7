import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-lb_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from lb_b200 import Lattice, native
name = sys.argv[1] if len(sys.argv) > 1 else "f32.strict.tma.v4.ty4.b6"
sim = Lattice(256, 64, 1.2, 1.01, 1.0, dtype=np.float32 if name.startswith("f32") else np.float64, math=name.split(".")[1])
sim.init_synthetic("pipe_ramp", amplitude=1e-3, seed=1)
sim.sync()
sim.set_variant(name)
try:
    sim.run(1)
    print("ran OK", float(sim.download("rho").mean()))
except Exception as e:
    print("ERR", e)
