"""lb_b200 -- B200-native D2Q9 lattice-Boltzmann engine behind the 2d-lb simulation-class API.

Layout of the package:
  native.py         ctypes binding of the C-ABI in include/lb_d2q9.h (csrc/liblb_d2q9.so)
  lattice.py        low-level lattice object (`Lattice.from_lattice`), one handle per slab
  dimensionless.py  drop-in for LB_D2Q9/dimensionless/opencl_dim.py (Pipe_Flow, Pipe_Flow_Cylinder,
                    Pipe_Flow_Obstacles)
  old_api.py        the pre-dimensionless constructor style of LB_D2Q9/OLD (Pipe_Flow_Obstacles(lx, ly, ...))
  slab.py           x-slab decomposition over torch.distributed, one process per GPU
There is no CPU fallback: every compute call goes through the CUDA library.
"""
from . import native  # noqa: F401
from .lattice import Lattice  # noqa: F401

__all__ = ["native", "Lattice"]
