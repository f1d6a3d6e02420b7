"""Import-compatibility shim: lets scripts written against the reference
(`from LB_D2Q9.dimensionless import opencl_dim as lb`) run unchanged on the CUDA engine
once `2d-lb_b200/` is on sys.path.  All code lives in `lb_b200`."""
