"""`LB_D2Q9.dimensionless.opencl_dim` of the reference, served by the B200 engine."""
from lb_b200.dimensionless import *  # noqa: F401,F403
from lb_b200.dimensionless import Pipe_Flow, Pipe_Flow_Cylinder, Pipe_Flow_Obstacles, get_divisible_global  # noqa: F401
