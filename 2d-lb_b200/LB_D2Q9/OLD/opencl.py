"""`LB_D2Q9.OLD.opencl` of the reference, served by the B200 engine (see lb_b200/old_api.py)."""
from lb_b200.old_api import (Pipe_Flow, Pipe_Flow_Obstacles, Pipe_Flow_PeriodicBC_VelocityInlet,  # noqa: F401
                             Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet)
