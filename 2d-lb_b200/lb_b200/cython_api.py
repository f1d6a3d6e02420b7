"""Drop-in for the reference's CPU classes `LB_D2Q9.dimensionless.cython_dim` (and, through
`scheme='cython_old'`, `LB_D2Q9.OLD.cython`), running on the GPU with the reference's own step
order and mixed-precision arithmetic (lb_cython.cuh).

Same constructor keywords, attributes and array conventions as cython_dim.pyx:
  f, feq : (9, nx, ny) float32, C-order        rho : (nx, ny) float32        u, v : (nx, ny) float64
and the same NumPy RNG consumption (randn(nx,ny) x2 in init_hydro, randn(nx,ny) in init_pop), so
that a seeded run reproduces the compiled reference BIT FOR BIT (tests/test_parity_gpu.py).
"""
import numpy as np

from . import dimensionless as _dim
from .lattice import Lattice

NUM_JUMPERS = 9


def _dev(a):
    """(9,nx,ny)/(nx,ny) host array -> device layout ([9,]ny,nx), contiguous."""
    a = np.asarray(a)
    return np.ascontiguousarray(a.transpose(0, 2, 1) if a.ndim == 3 else a.T)


def _host(a):
    """device layout -> the Cython classes' (9,nx,ny)/(nx,ny) C-order arrays."""
    return np.ascontiguousarray(a.transpose(0, 2, 1) if a.ndim == 3 else a.T)


class Pipe_Flow(_dim.Pipe_Flow):
    """cython_dim.Pipe_Flow (cython_dim.pyx:31-396)."""

    _scheme = "cython"

    def __init__(self, diameter=None, rho=None, viscosity=None, pressure_grad=1., pipe_length=None,
                 N=100, time_prefactor=1., device=0, verbose=True):
        super(Pipe_Flow, self).__init__(diameter=diameter, rho=rho, viscosity=viscosity, pressure_grad=pressure_grad,
                                        pipe_length=pipe_length, N=N, time_prefactor=time_prefactor,
                                        device=device, verbose=verbose, units="cython", dtype=np.float32)

    def init_cuda(self):
        self.sim = Lattice(self.nx, self.ny, self.omega, self.inlet_rho, self.outlet_rho, bc="pipe",
                           dtype=np.float32, device=self._device, scheme=self._scheme)

    def init_hydro(self):
        """cython_dim.pyx:130-157"""
        nx, ny = self.nx, self.ny
        self._set_boundary_densities()
        self._say('inlet rho:', self.inlet_rho)
        self._say('outlet rho:', self.outlet_rho)
        rho = np.ones((nx, ny), dtype=np.float32)
        rho[0, :] = self.inlet_rho
        rho[self.lx, :] = self.outlet_rho
        for i in range(rho.shape[0]):
            rho[i, :] = self.inlet_rho - i * (self.inlet_rho - self.outlet_rho) / float(rho.shape[0])
        u = .0 * np.random.randn(nx, ny)
        v = .0 * np.random.randn(nx, ny)
        self._upload_hydro(rho, u, v)

    def _upload_hydro(self, rho, u, v):
        self.sim.upload_moments(_dev(rho), _dev(u), _dev(v))

    def init_pop(self):
        """cython_dim.pyx:191-202: one perturbation per node, shared by the nine populations."""
        f = _host(self.sim.download("feq"))
        amplitude = .001
        perturb = (1. + amplitude * np.random.randn(self.nx, self.ny))
        f *= perturb
        self.sim.upload_f(_dev(f))

    # the Cython classes keep their fields as attributes
    f = property(lambda self: _host(self.sim.download("f")))
    feq = property(lambda self: _host(self.sim.download("feq")))
    rho = property(lambda self: _host(self.sim.download("rho")))
    u = property(lambda self: _host(self.sim.download("u")))
    v = property(lambda self: _host(self.sim.download("v")))

    # the single steps of cython_dim.pyx:204-344, individually callable like the reference's methods
    # (run() is the same sequence fused into one kernel per step)
    def move_bcs(self):
        self.sim.move_bcs()

    def move(self):
        self.sim.move()

    def update_hydro(self):
        self.sim.update_hydro()

    def collide_particles(self):
        self.sim.collide_particles()

    def get_fields(self):
        """cython_dim.pyx:361-372"""
        return {'f': self.f, 'u': self.u, 'v': self.v, 'rho': self.rho, 'feq': self.feq}


class Pipe_Flow_Cylinder(Pipe_Flow):
    """cython_dim.Pipe_Flow_Cylinder (cython_dim.pyx:398-513): u, v are zeroed inside the obstacle after
    every moment update."""

    def __init__(self, cylinder_center=None, cylinder_radius=None, **kwargs):
        assert (cylinder_center is not None)
        assert (cylinder_radius is not None)
        self.phys_cylinder_center = cylinder_center
        self.phys_cylinder_radius = cylinder_radius
        self.obstacle_mask = None
        super(Pipe_Flow_Cylinder, self).__init__(**kwargs)
        self.obstacle_pixels = np.where(self.obstacle_mask)

    def set_characteristic_length_time(self):
        """cython_dim.pyx:405-408"""
        self.L = self.phys_cylinder_radius
        self.T = (8 * self.phys_rho * self.phys_visc * self.L) / (np.abs(self.phys_pressure_grad) * self.phys_diameter ** 2)

    def initialize_grid_dims(self):
        """cython_dim.pyx:410-431"""
        _dim.Pipe_Flow_Cylinder.initialize_grid_dims(self)
        self.obstacle_mask = np.asfortranarray(self.obstacle_mask_host.astype(bool))

    def init_hydro(self):
        super(Pipe_Flow_Cylinder, self).init_hydro()
        self.sim.set_mask(np.asarray(self.obstacle_mask).T)
        self.sim.zero_velocity_in_obstacle()
