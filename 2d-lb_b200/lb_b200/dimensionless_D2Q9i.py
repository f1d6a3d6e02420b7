"""Drop-in for `LB_D2Q9.dimensionless.opencl_dim_D2Q9i` (the incompressible D2Q9i variant).

Host differences from opencl_dim.py, all taken from opencl_dim_D2Q9i.py: the Reynolds-number
parameterisation of the Cython module (T = 8 rho nu/(|grad p| L), omega = 1/(nu_lb/cs2 + 0.5),
pressure drop scaled by the non-dimensional gradient: :98-120, :180, :253-255, :440), no
`use_interop` keyword (:65), kernels from D2Q9i.cl (:228), and u, v zeroed inside the cylinder
after EVERY moment update (:490-499).  Device side: `model='d2q9i'` of the fused kernel.
"""
import numpy as np

from . import dimensionless as _dim
from .lattice import Lattice


class Pipe_Flow(_dim.Pipe_Flow):
    def __init__(self, diameter=None, rho=None, viscosity=None, pressure_grad=None, pipe_length=None,
                 N=200, time_prefactor=1., two_d_local_size=(32, 32), three_d_local_size=(32, 32, 1),
                 dtype=np.float32, math="strict", device=0, verbose=True):
        super(Pipe_Flow, self).__init__(diameter=diameter, rho=rho, viscosity=viscosity, pressure_grad=pressure_grad,
                                        pipe_length=pipe_length, N=N, time_prefactor=time_prefactor,
                                        two_d_local_size=two_d_local_size, three_d_local_size=three_d_local_size,
                                        dtype=dtype, math=math, device=device, verbose=verbose, units="cython")

    def init_cuda(self):
        self.sim = Lattice(self.nx, self.ny, self.omega, self.inlet_rho, self.outlet_rho, bc="pipe",
                           dtype=self.dtype, math=self._math, device=self._device, model="d2q9i",
                           zero_obstacle_velocity=self._zero_vel)


class Pipe_Flow_Cylinder(Pipe_Flow):
    """opencl_dim_D2Q9i.py:426-510"""

    def __init__(self, cylinder_center=None, cylinder_radius=None, **kwargs):
        assert (cylinder_center is not None)
        assert (cylinder_radius is not None)
        self.phys_cylinder_center = cylinder_center
        self.phys_cylinder_radius = cylinder_radius
        self.obstacle_mask_host = None
        self._zero_vel = True                       # update_hydro override, :494-499
        super(Pipe_Flow_Cylinder, self).__init__(**kwargs)

    def set_characteristic_length_time(self):
        self.L = self.phys_cylinder_radius
        self.T = (8 * self.phys_rho * self.phys_visc * self.L) / (np.abs(self.phys_pressure_grad) * self.phys_diameter ** 2)

    initialize_grid_dims = _dim.Pipe_Flow_Cylinder.initialize_grid_dims

    def init_hydro(self):
        super(Pipe_Flow_Cylinder, self).init_hydro()
        self.sim.set_mask(np.asarray(self.obstacle_mask_host).T)
        self.sim.zero_velocity_in_obstacle()
