"""Drop-in for the reference's `LB_D2Q9.OLD.cython` classes (OLD/cython.pyx), running on the GPU with
that module's own step order and arithmetic (`scheme='cython_old'`, lb_cython.cuh):

  Pipe_Flow(omega, lx, ly, dr, dt, deltaP)                              OLD/cython.pyx:31-265
  Pipe_Flow_Obstacles(obstacle_mask=..., **kw)                          :425-486
  Pipe_Flow_PeriodicBC_VelocityInlet(u_w, **kw)                         :268-360   
  Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet(obstacle_mask=..., **kw) :362-423

Arrays follow the Cython classes: f, feq (9, nx, ny) float32; rho (nx, ny) float32; u, v (nx, ny)
float64.  A seeded run reproduces the compiled reference bit for bit (tests/test_parity_gpu.py).
"""
import numpy as np

from .cython_api import _dev, _host
from .lattice import Lattice, cs2

NUM_JUMPERS = 9


class Pipe_Flow(object):
    _bc = "pipe"

    def __init__(self, omega=.99, lx=400, ly=400, dr=1., dt=1., deltaP=-.1, device=0):
        self.lx, self.ly = lx, ly
        self.omega = omega
        self.dr, self.dt, self.deltaP = dr, dt, deltaP
        self.nx, self.ny = self.lx + 1, self.ly + 1
        self.inlet_rho = 1.
        self.outlet_rho = self.deltaP / cs2 + self.inlet_rho          # OLD/cython.pyx:47-51
        self.sim = Lattice(self.nx, self.ny, self.omega, self.inlet_rho, float(self.outlet_rho), bc=self._bc,
                           dtype=np.float32, device=device, scheme="cython_old",
                           u_west=getattr(self, "u_w", 0.0), u_east=getattr(self, "u_e", 0.0))
        self.init_hydro()
        self.update_feq()
        self.init_pop()

    def init_hydro(self):
        """OLD/cython.pyx:83-95"""
        nx, ny = self.nx, self.ny
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

    def update_feq(self):
        self.sim.update_feq()

    def init_pop(self):
        """OLD/cython.pyx:241-250 (amplitude 0, RNG still consumed)"""
        f = _host(self.sim.download("feq"))
        amplitude = .00
        perturb = (1. + amplitude * np.random.randn(self.nx, self.ny))
        f *= perturb
        self.sim.upload_f(_dev(f))

    def run(self, num_iterations):
        self.sim.run(int(num_iterations))

    # the single steps (OLD/cython.pyx:97-256), individually callable like the reference's methods
    def move_bcs(self):
        self.sim.move_bcs()

    def move(self):
        self.sim.move()

    def update_hydro(self):
        self.sim.update_hydro()

    def collide_particles(self):
        self.sim.collide_particles()

    f = property(lambda self: _host(self.sim.download("f")))
    feq = property(lambda self: _host(self.sim.download("feq")))
    rho = property(lambda self: _host(self.sim.download("rho")))
    u = property(lambda self: _host(self.sim.download("u")))
    v = property(lambda self: _host(self.sim.download("v")))


class _WithObstacles(object):
    def _install_mask(self, obstacle_mask):
        self.obstacle_mask = np.asarray(obstacle_mask, dtype=bool)
        self.obstacle_pixels = np.where(self.obstacle_mask)

    def _apply_mask(self):
        if self.obstacle_mask.shape != (self.nx, self.ny):
            raise ValueError(f"obstacle_mask must have shape (nx, ny) = {(self.nx, self.ny)}")
        self.sim.set_mask(self.obstacle_mask.T)
        self.sim.zero_velocity_in_obstacle()


class Pipe_Flow_Obstacles(_WithObstacles, Pipe_Flow):
    def __init__(self, *args, obstacle_mask=None, **kwargs):
        self._install_mask(obstacle_mask)
        Pipe_Flow.__init__(self, *args, **kwargs)

    def init_hydro(self):
        Pipe_Flow.init_hydro(self)
        self._apply_mask()


class Pipe_Flow_PeriodicBC_VelocityInlet(Pipe_Flow):
    _bc = "velocity_yperiodic"

    def __init__(self, u_w=0.1, **kwargs):
        self.u_w = u_w
        self.u_e = u_w
        Pipe_Flow.__init__(self, **kwargs)

    def init_hydro(self):
        """OLD/cython.pyx:319-328: rho = 1, u = u_w everywhere, v = 0 (no RNG)."""
        nx, ny = self.nx, self.ny
        rho = np.ones((nx, ny), dtype=np.float32)
        u = np.zeros((nx, ny))
        v = np.zeros((nx, ny))
        u[:, :] = self.u_w
        self._upload_hydro(rho, u, v)


class Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet(_WithObstacles, Pipe_Flow_PeriodicBC_VelocityInlet):
    def __init__(self, *args, obstacle_mask=None, **kwargs):
        self._install_mask(obstacle_mask)
        Pipe_Flow_PeriodicBC_VelocityInlet.__init__(self, *args, **kwargs)

    def init_hydro(self):
        Pipe_Flow_PeriodicBC_VelocityInlet.init_hydro(self)
        self._apply_mask()
