"""The pre-dimensionless constructor style of the reference (LB_D2Q9/OLD/opencl.py).

`north_star` names "Pipe_Flow_Obstacles style construction": in the reference the live class with
that name takes lattice parameters directly -- Pipe_Flow_Obstacles(obstacle_mask=..., omega=...,
lx=..., ly=..., dr=..., dt=..., deltaP=...) (OLD/opencl.py:44-62, :373-415).  The shipped
OLD/opencl.py cannot run (it opens OLD/D2Q9.cl, which does not exist -- SURVEY.md F13), so
these classes keep its constructor, attributes and `get_fields_on_cpu()` but step with the
working kernel order of dimensionless/opencl_dim.py:372-387 (stream, then BCs).
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
