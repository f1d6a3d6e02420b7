"""`LB_D2Q9.dimensionless.cython_dim` of the reference, served by the B200 engine (Cython-order scheme)."""
from lb_b200.cython_api import Pipe_Flow, Pipe_Flow_Cylinder  # noqa: F401
