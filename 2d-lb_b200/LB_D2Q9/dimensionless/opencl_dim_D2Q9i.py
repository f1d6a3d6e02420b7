"""`LB_D2Q9.dimensionless.opencl_dim_D2Q9i` of the reference, served by the B200 engine."""
from lb_b200.dimensionless_D2Q9i import Pipe_Flow, Pipe_Flow_Cylinder  # noqa: F401
