"""`LB_D2Q9.OLD.opencl` of the reference (constructor style only), served by the B200 engine."""
from lb_b200.old_api import Pipe_Flow, Pipe_Flow_Obstacles  # noqa: F401
