"""`LB_D2Q9.OLD.cython` of the reference, served by the B200 engine (scheme 'cython_old')."""
from lb_b200.old_cython_api import (Pipe_Flow, Pipe_Flow_Obstacles, Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet,  # noqa: F401
                                    Pipe_Flow_PeriodicBC_VelocityInlet)
