"""The pre-dimensionless constructor style of the reference (LB_D2Q9/OLD/opencl.py).

`north_star` names "Pipe_Flow_Obstacles style construction": in the reference the live class with
that name takes lattice parameters directly -- Pipe_Flow_Obstacles(obstacle_mask=..., omega=...,
lx=..., ly=..., dr=..., dt=..., deltaP=...) (OLD/opencl.py:44-62, :373-415).  The shipped
OLD/opencl.py cannot run as is (it opens OLD/D2Q9.cl, which does not exist -- SURVEY.md F13).
Pointed at LB_D2Q9/D2Q9.cl and executed on the CPU emulation the tests use, its two families
behave differently (tests/test_opencl_reference.py, tests/golden/oldcl_*.npz):

  * Pipe_Flow / Pipe_Flow_Obstacles (pressure-driven): the OLD step order applies D2Q9.cl's `move_bcs`
    -- written for freshly streamed populations -- before streaming, and the run diverges (NaN within
    ~250 steps at omega = 1, deltaP = -1e-3; 1e6 within 5 steps behind an obstacle).  These classes
    therefore keep the constructor, attributes and `get_fields_on_cpu()` but step with the working
    kernel order of dimensionless/opencl_dim.py:372-387 (stream, then BCs).
  * Pipe_Flow_PeriodicBC_VelocityInlet / Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet (:281-371): stable,
    and the only callers of D2Q9.cl:263-374.  Reproduced exactly -- step order, the populations `move`
    never writes, the velocity entries update_hydro never rewrites -- by scheme 'opencl_old'
    (csrc/lb_oldcl.cuh): a seeded run equals the reference's own kernels bit for bit.
"""
import numpy as np

from .lattice import Lattice, cs2

NUM_JUMPERS = 9


class Pipe_Flow(object):
    def __init__(self, omega=.99, lx=400, ly=400, dr=1., dt=1., deltaP=-.1,
                 two_d_local_size=(32, 32), three_d_local_size=(32, 32, 1),
                 dtype=np.float32, math="strict", device=0):
        self.lx, self.ly = lx, ly
        self.omega = np.float32(omega)                       # OLD/opencl.py:50
        self.dr, self.dt, self.deltaP = np.float32(dr), np.float32(dt), np.float32(deltaP)
        self.nx, self.ny = self.lx + 1, self.ly + 1
        self.inlet_rho = 1.
        self.outlet_rho = self.deltaP / cs2 + self.inlet_rho  # deltaP < 0, OLD/opencl.py:60-62
        self.two_d_local_size, self.three_d_local_size = two_d_local_size, three_d_local_size
        self.dtype = np.dtype(dtype)
        if not hasattr(self, "_zero_vel"):
            self._zero_vel = False
        self.sim = Lattice(self.nx, self.ny, float(self.omega), float(self.inlet_rho), float(self.outlet_rho),
                           bc="pipe", dtype=self.dtype, math=math, device=device,
                           zero_obstacle_velocity=self._zero_vel)
        self.init_hydro()
        self.update_feq()
        self.init_pop()

    def init_hydro(self):
        """OLD/opencl.py:166-185"""
        nx, ny = self.nx, self.ny
        rho_host = np.ones((nx, ny), dtype=np.float32, order='F')
        for i in range(nx):
            rho_host[i, :] = self.inlet_rho - i * (self.inlet_rho - self.outlet_rho) / float(nx)
        u_host = (.0 * np.random.randn(nx, ny)).astype(np.float32, order='F')
        v_host = (.0 * np.random.randn(nx, ny)).astype(np.float32, order='F')
        self.sim.upload_moments(rho_host.T, u_host.T, v_host.T)

    def update_feq(self):
        self.sim.update_feq()

    def init_pop(self):
        """OLD/opencl.py:204-222 (noise amplitude 0, but the RNG is still consumed)"""
        f = np.zeros((self.nx, self.ny, NUM_JUMPERS), dtype=self.dtype, order='F')
        self.sim.download("feq", out=f.T)
        amplitude = .00
        f *= (1. + amplitude * np.random.randn(self.nx, self.ny, NUM_JUMPERS))
        self.sim.upload_f(f.T)

    def move_bcs(self):
        self.sim.move_bcs()

    def move(self):
        self.sim.move()

    def update_hydro(self):
        self.sim.update_hydro()

    def collide_particles(self):
        self.sim.collide_particles()

    def run(self, num_iterations):
        self.sim.run(int(num_iterations))

    def get_fields_on_cpu(self):
        """OLD/opencl.py:257-279"""
        out = {}
        for name in ('f', 'feq'):
            a = np.zeros((self.nx, self.ny, NUM_JUMPERS), dtype=self.dtype, order='F')
            self.sim.download(name, out=a.T)
            out[name] = a
        for name in ('u', 'v', 'rho'):
            a = np.zeros((self.nx, self.ny), dtype=self.dtype, order='F')
            self.sim.download(name, out=a.T)
            out[name] = a
        return out

    get_fields = get_fields_on_cpu


class Pipe_Flow_Obstacles(Pipe_Flow):
    """OLD/opencl.py:373-415"""

    def __init__(self, obstacle_mask=None, **kwargs):
        assert (obstacle_mask is not None)
        assert (np.sum(obstacle_mask) != 0)
        self.obstacle_mask_host = np.asfortranarray(obstacle_mask).astype(np.int32)
        self._zero_vel = True
        super(Pipe_Flow_Obstacles, self).__init__(**kwargs)

    def init_hydro(self):
        super(Pipe_Flow_Obstacles, self).init_hydro()
        if self.obstacle_mask_host.shape != (self.nx, self.ny):
            raise ValueError(f"obstacle_mask must have shape (nx, ny) = {(self.nx, self.ny)}")
        self.sim.set_mask(np.asarray(self.obstacle_mask_host).T)
        self.sim.zero_velocity_in_obstacle()


class Pipe_Flow_PeriodicBC_VelocityInlet(Pipe_Flow):
    """OLD/opencl.py:281-327 -- imposed x-velocity u_w at inlet and outlet, rows 0 and ny-1 periodic."""

    def __init__(self, u_w=0.1, omega=.99, lx=400, ly=400, dr=1., dt=1., deltaP=-.1,
                 two_d_local_size=(32, 32), three_d_local_size=(32, 32, 1), device=0):
        self.u_w = u_w
        self.u_e = u_w
        self.lx, self.ly = lx, ly
        self.omega = np.float32(omega)
        self.dr, self.dt, self.deltaP = np.float32(dr), np.float32(dt), np.float32(deltaP)
        self.nx, self.ny = self.lx + 1, self.ly + 1
        self.inlet_rho = 1.
        self.outlet_rho = self.deltaP / cs2 + self.inlet_rho
        self.two_d_local_size, self.three_d_local_size = two_d_local_size, three_d_local_size
        self.dtype = np.dtype(np.float32)
        # np.float32(self.u_w), np.float32(self.u_e) are what the kernels receive (:293-294, :323-324)
        self.sim = Lattice(self.nx, self.ny, float(self.omega), bc="velocity_yperiodic", dtype=np.float32,
                           device=device, scheme="opencl_old", u_west=float(np.float32(self.u_w)),
                           u_east=float(np.float32(self.u_e)))
        self.init_hydro()
        self.update_feq()
        self.init_pop()

    def init_hydro(self):
        """OLD/opencl.py:299-316: rho = 1, u = u_w everywhere, v = 0 (no RNG)."""
        nx, ny = self.nx, self.ny
        rho_host = np.ones((nx, ny), dtype=np.float32, order='F')
        u_host = (np.ones((nx, ny)) * self.u_w).astype(np.float32, order='F')
        v_host = np.zeros((nx, ny)).astype(np.float32, order='F')
        self.sim.upload_moments(rho_host.T, u_host.T, v_host.T)



class Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet(Pipe_Flow_PeriodicBC_VelocityInlet):
    """OLD/opencl.py:329-371"""

    def __init__(self, obstacle_mask=None, **kwargs):
        assert (obstacle_mask is not None)
        assert (np.sum(obstacle_mask) != 0)
        obstacle_mask = np.asfortranarray(obstacle_mask)
        self.obstacle_mask_host = obstacle_mask.astype(np.int32)
        super(Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet, self).__init__(**kwargs)

    def init_hydro(self):
        super(Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet, self).init_hydro()
        if self.obstacle_mask_host.shape != (self.nx, self.ny):
            raise ValueError(f"obstacle_mask must have shape (nx, ny) = {(self.nx, self.ny)}")
        self.sim.set_mask(np.asarray(self.obstacle_mask_host).T)
        self.sim.zero_velocity_in_obstacle()
