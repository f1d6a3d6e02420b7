"""Drop-in for the reference's `LB_D2Q9.dimensionless.opencl_dim` simulation classes.

Same class names, constructor keywords, attributes, method names and return layouts as
LB_D2Q9/dimensionless/opencl_dim.py; everything underneath (`cl.Context`, `cl.Buffer`,
`kernels.<name>(...).wait()`, `cl.enqueue_copy`) is replaced by calls into the CUDA library
through `lb_b200.lattice.Lattice`.  Host-side parameter algebra and NumPy RNG consumption
follow the reference line by line so that the same seed gives the same initial populations:

  Pipe_Flow.__init__            opencl_dim.py:64-178     nondimensionalisation :103-120
  set_characteristic_length_time  :180-189               initialize_grid_dims  :191-201
  init_hydro  :258-293          update_feq :295-306      init_pop :308-327
  run :372-387                  get_fields :390-415      get_nondim_fields / get_physical_fields :417-438
  Pipe_Flow_Cylinder            :441-518                 Pipe_Flow_Obstacles (template) :616-657

Extensions (keyword-only, defaults reproduce the reference): dtype, math, device, verbose,
zero_obstacle_velocity_each_step, devices=[...] (x-slab decomposition over several GPUs of this
process, peer-memory halos; the reference is single-device), and units='opencl' | 'cython'.  The reference's OpenCL and
Cython modules map the same physical inputs to different lattice parameters:
units='cython' selects cython_dim.pyx's algebra (T = 8 rho nu/(|grad p| L), Reynolds-number based
omega, pressure drop scaled by the non-dimensional gradient; cython_dim.pyx:70,85-88,115-116,
139-144,407-408) -- the one the published benchmark notebook's printouts were made with.
"""
import numpy as np

from . import draw
from .lattice import Lattice, LocalSlabs, cs2

NUM_JUMPERS = 9

# D2Q9 parameters as module attributes, like the reference (opencl_dim.py:22-36)
w = np.array([4. / 9., 1. / 9., 1. / 9., 1. / 9., 1. / 9., 1. / 36., 1. / 36., 1. / 36., 1. / 36.],
             order='F', dtype=np.float32)
cx = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], order='F', dtype=np.int32)
cy = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], order='F', dtype=np.int32)


def get_divisible_global(global_size, local_size):
    """Smallest multiple of local_size covering global_size (opencl_dim.py:39-56).  Kept for
    scripts that print it; launch geometry is chosen by the CUDA library."""
    return tuple(g if g % l == 0 else g + l - g % l for g, l in zip(global_size, local_size))


class Pipe_Flow(object):
    """Pressure-driven flow between two plates on the D2Q9 lattice (opencl_dim.py:58-438)."""

    def __init__(self, diameter=None, rho=None, viscosity=None, pressure_grad=None, pipe_length=None,
                 N=200, time_prefactor=1.,
                 two_d_local_size=(32, 32), three_d_local_size=(32, 32, 1), use_interop=False,
                 dtype=np.float32, math="strict", device=0, verbose=True,
                 zero_obstacle_velocity_each_step=None, units="opencl", devices=None):
        self._devices = list(devices) if devices else None
        if units not in ("opencl", "cython"):
            raise ValueError("units must be 'opencl' or 'cython'")
        self._units = units
        self._verbose = verbose
        self.dtype = np.dtype(dtype)
        self._math = math
        self._device = device
        if zero_obstacle_velocity_each_step is not None:
            self._zero_vel = bool(zero_obstacle_velocity_each_step)
        elif not hasattr(self, "_zero_vel"):
            self._zero_vel = False

        # Physical units (opencl_dim.py:85-91)
        self.phys_diameter = diameter
        self.phys_rho = rho
        self.phys_visc = viscosity
        self.phys_pressure_grad = pressure_grad
        self.phys_pressure_grad_div_rho = self.phys_pressure_grad / self.phys_rho
        self.phys_pipe_length = pipe_length
        self.use_interop = use_interop          # accepted, ignored (no GL interop on a headless GPU)

        self.L = None
        self.T = None
        self.set_characteristic_length_time()
        self._say('Characteristic L:', self.L)
        self._say('Characteristic T:', self.T)

        if self._units == "opencl":
            self.W = (np.abs(self.phys_pressure_grad_div_rho) * self.L * self.T) / self.phys_visc
            self._say('Weinstein number:', self.W)
        else:
            self.Re = self.L ** 2 / (self.phys_visc * self.T ** 2)        # cython_dim.pyx:70
            self._say('Reynolds number:', self.Re)

        self.N = N
        self.delta_x = 1. / N
        self.delta_t = time_prefactor * self.delta_x ** 2
        self.ulb = self.delta_t / self.delta_x
        self._say('u_lb:', self.ulb)

        if self._units == "opencl":
            self.lb_viscosity = (self.delta_t / self.delta_x ** 2) * (1. / self.W)
            self.omega = (3 * self.lb_viscosity + 0.5) ** -1.
        else:                                                             # cython_dim.pyx:85-88
            self.lb_viscosity = (self.delta_t / self.delta_x ** 2) * (1. / self.Re)
            self.omega = (self.lb_viscosity / cs2 + 0.5) ** -1.
        self._say('omega', self.omega)
        assert self.omega < 2.

        self.lx = None
        self.ly = None
        self.nx = None
        self.ny = None
        self.initialize_grid_dims()

        # work-group sizes are accepted for compatibility and reported like the reference does
        self.two_d_local_size = two_d_local_size
        self.three_d_local_size = three_d_local_size
        self.two_d_global_size = get_divisible_global((self.nx, self.ny), self.two_d_local_size)
        self.three_d_global_size = get_divisible_global((self.nx, self.ny, 9), self.three_d_local_size)
        self._say('2d global:', self.two_d_global_size)
        self._say('2d local:', self.two_d_local_size)
        self._say('3d global:', self.three_d_global_size)
        self._say('3d local:', self.three_d_local_size)

        # inlet/outlet densities are needed to create the device lattice (opencl_dim.py:270-274)
        self.inlet_rho = None
        self.outlet_rho = None
        self._set_boundary_densities()

        self.sim = None            # the CUDA lattice (replaces context/queue/kernels/buffers)
        self.init_cuda()

        self.init_hydro()
        self.update_feq()
        self.init_pop()

    # -- helpers ------------------------------------------------------------------------
    def _say(self, *args):
        if self._verbose:
            print(*args)

    def set_characteristic_length_time(self):
        """opencl_dim.py:180-189 (units='cython': cython_dim.pyx:115-116)"""
        self.L = self.phys_diameter
        if self._units == "cython":
            self.T = (8 * self.phys_rho * self.phys_visc) / (np.abs(self.phys_pressure_grad) * self.L)
            return
        zeta = np.abs(self.phys_pressure_grad) / self.phys_rho
        self.T = np.sqrt(self.phys_diameter / zeta)

    def initialize_grid_dims(self):
        """opencl_dim.py:191-201"""
        self.lx = int(np.ceil((self.phys_pipe_length / self.L) * self.N))
        self.ly = self.N
        self.nx = self.lx + 1
        self.ny = self.ly + 1

    def _set_boundary_densities(self):
        """opencl_dim.py:268-274 (units='cython': cython_dim.pyx:139-144)"""
        nondim_gradP = 1.
        if self._units == "cython":
            nondim_gradP = (self.T ** 2 / (self.phys_rho * self.L)) * self.phys_pressure_grad
        delta_rho = self.nx * (self.delta_t ** 2 / self.delta_x) * (1. / cs2) * nondim_gradP
        self.outlet_rho = 1.
        self.inlet_rho = 1. + np.abs(delta_rho)

    def init_cuda(self):
        """Replaces init_opencl + allocate_constants (opencl_dim.py:203-255)."""
        kw = dict(omega=self.omega, inlet_rho=self.inlet_rho, outlet_rho=self.outlet_rho, bc="pipe",
                  dtype=self.dtype, math=self._math, zero_obstacle_velocity=self._zero_vel)
        if self._devices and len(self._devices) > 1:
            self.sim = LocalSlabs(self.nx, self.ny, devices=self._devices, **kw)      # one x-slab per GPU
        else:
            self.sim = Lattice(self.nx, self.ny, device=self._devices[0] if self._devices else self._device, **kw)

    # -- initialisation (individually callable) -------------------------------
    def init_hydro(self):
        """opencl_dim.py:258-293"""
        nx, ny = self.nx, self.ny
        self._set_boundary_densities()
        self._say('inlet rho:', self.inlet_rho)
        self._say('outlet rho:', self.outlet_rho)

        rho_host = self.inlet_rho * np.ones((nx, ny), dtype=np.float32, order='F')
        rho_host[0, :] = self.inlet_rho
        rho_host[self.lx, :] = self.outlet_rho
        for i in range(rho_host.shape[0]):
            rho_host[i, :] = self.inlet_rho - i * (self.inlet_rho - self.outlet_rho) / float(rho_host.shape[0])

        u_host = .0 * np.random.randn(nx, ny)          # consumes the RNG exactly like the reference
        u_host = u_host.astype(np.float32, order='F')
        v_host = .0 * np.random.randn(nx, ny)
        v_host = v_host.astype(np.float32, order='F')

        # (nx, ny) Fortran-order == device [ny][nx]; .T is a zero-copy C-order view
        self.sim.upload_moments(rho_host.T, u_host.T, v_host.T)

    def update_feq(self):
        """opencl_dim.py:295-306"""
        self.sim.update_feq()

    def init_pop(self):
        """opencl_dim.py:308-327: f = feq * (1 + 0.001 * randn(nx, ny, 9)), uploaded to both buffers."""
        nx, ny = self.nx, self.ny
        f = np.zeros((nx, ny, NUM_JUMPERS), dtype=self.dtype, order='F')
        self.sim.download("feq", out=f.T)
        amplitude = .001
        perturb = (1. + amplitude * np.random.randn(nx, ny, NUM_JUMPERS))
        f *= perturb
        self.sim.upload_f(f.T)

    # -- single steps ---------------------------------------------------------------------
    def move_bcs(self):
        """opencl_dim.py:329-337 (and :510-518 when a mask is present)"""
        self.sim.move_bcs()

    def move(self):
        """opencl_dim.py:339-353"""
        self.sim.move()

    def update_hydro(self):
        """opencl_dim.py:355-362"""
        self.sim.update_hydro()

    def collide_particles(self):
        """opencl_dim.py:364-370"""
        self.sim.collide_particles()

    def run(self, num_iterations):
        """opencl_dim.py:372-387 -- here ONE fused launch per iteration, one host sync at the end."""
        self.sim.run(int(num_iterations))

    # -- readback -------------------------------------------------------------------------
    def get_fields(self):
        """opencl_dim.py:390-415: dict of Fortran-order host arrays, f/feq (nx,ny,9), u/v/rho (nx,ny)."""
        nx, ny = self.nx, self.ny
        results = {}
        for name in ('f', 'feq'):
            a = np.zeros((nx, ny, NUM_JUMPERS), dtype=self.dtype, order='F')
            self.sim.download(name, out=a.T)
            results[name] = a
        for name in ('u', 'v', 'rho'):
            a = np.zeros((nx, ny), dtype=self.dtype, order='F')
            self.sim.download(name, out=a.T)
            results[name] = a
        return results

    # -- device handles (opencl_dim.py:165-176, :231: `sim.queue`, `sim.u`, `sim.v`, `sim.rho`, `sim.f`, `sim.feq`
    #    are pyopencl objects there; the visualiser reads `sim.u.get()` every frame, field_visualizer.py:146-157)
    @property
    def queue(self):
        return _Queue(self.sim)

    f = property(lambda self: _DeviceField(self, 'f'))
    feq = property(lambda self: _DeviceField(self, 'feq'))
    rho = property(lambda self: _DeviceField(self, 'rho'))
    u = property(lambda self: _DeviceField(self, 'u'))
    v = property(lambda self: _DeviceField(self, 'v'))

    def get_nondim_fields(self):
        """opencl_dim.py:417-426"""
        fields = self.get_fields()
        fields['u'] *= self.delta_x / self.delta_t
        fields['v'] *= self.delta_x / self.delta_t
        return fields

    def get_physical_fields(self):
        """opencl_dim.py:428-438"""
        fields = self.get_nondim_fields()
        fields['u'] *= (self.L / self.T)
        fields['v'] *= (self.L / self.T)
        return fields


class _Queue:
    """Stand-in for the `cl.CommandQueue` attribute: `finish()` waits for the device."""

    def __init__(self, sim):
        self._sim = sim

    def finish(self):
        self._sim.sync()

    flush = finish


class _DeviceField:
    """Stand-in for a `pyopencl.array.Array` / `cl.Buffer` attribute of the simulation: `.get()` returns the
    host copy in the reference's shape and order ((nx, ny[, 9]), Fortran), `.ptr` / `.pitch` the device
    address and row pitch in elements (single-slab lattices) for zero-copy consumers."""

    def __init__(self, owner, name):
        self._owner, self.name = owner, name

    @property
    def shape(self):
        o = self._owner
        return (o.nx, o.ny, NUM_JUMPERS) if self.name in ('f', 'feq') else (o.nx, o.ny)

    @property
    def dtype(self):
        return np.dtype(self._owner.sim.field_dtype(self.name))

    def get(self):
        a = np.zeros(self.shape, dtype=self.dtype, order='F')
        self._owner.sim.download(self.name, out=a.T)
        return a

    @property
    def ptr(self):
        return self._owner.sim.device_ptr(self.name)[0]

    @property
    def pitch(self):
        return self._owner.sim.device_ptr(self.name)[1]


class Pipe_Flow_Cylinder(Pipe_Flow):
    """Flow around a cylinder (opencl_dim.py:441-518).  `obstacle_mask_host` may be overwritten
    and `init_hydro(); update_feq(); init_pop()` re-run to simulate arbitrary obstacles, as
    docs/cs205_movie.ipynb does."""

    def __init__(self, cylinder_center=None, cylinder_radius=None, **kwargs):
        assert (cylinder_center is not None)
        assert (cylinder_radius is not None)
        self.phys_cylinder_center = cylinder_center
        self.phys_cylinder_radius = cylinder_radius
        self.obstacle_mask_host = None
        super(Pipe_Flow_Cylinder, self).__init__(**kwargs)

    def set_characteristic_length_time(self):
        """opencl_dim.py:448-457 (units='cython': cython_dim.pyx:407-408)"""
        self.L = self.phys_cylinder_radius
        if self._units == "cython":
            self.T = (8 * self.phys_rho * self.phys_visc * self.L) / (np.abs(self.phys_pressure_grad) * self.phys_diameter ** 2)
            return
        zeta = np.abs(self.phys_pressure_grad) / self.phys_rho
        self.T = np.sqrt(self.phys_cylinder_radius / zeta)

    def initialize_grid_dims(self):
        """opencl_dim.py:459-475"""
        self.lx = int(np.ceil((self.phys_pipe_length / self.L) * self.N))
        self.ly = int(np.ceil((self.phys_diameter / self.L) * self.N))
        self.nx = self.lx + 1
        self.ny = self.ly + 1
        self.obstacle_mask_host = np.zeros((self.nx, self.ny), dtype=np.int32, order='F')
        x_cylinder = self.N * self.phys_cylinder_center[0] / self.L
        y_cylinder = self.N * self.phys_cylinder_center[1] / self.L
        circle = draw.circle(x_cylinder, y_cylinder, self.N)
        self.obstacle_mask_host[circle[0], circle[1]] = 1

    def init_hydro(self):
        """opencl_dim.py:495-508: upload the mask, zero u,v inside it (at initialisation only)."""
        super(Pipe_Flow_Cylinder, self).init_hydro()
        self.sim.set_mask(np.asarray(self.obstacle_mask_host).T)
        self.sim.zero_velocity_in_obstacle()


class Pipe_Flow_Obstacles(Pipe_Flow):
    """Pipe flow around an arbitrary obstacle mask: the class the reference keeps as a commented
    template in opencl_dim.py:616-657 (live in LB_D2Q9/OLD/opencl.py:373-415).  Unlike
    Pipe_Flow_Cylinder it zeroes u,v inside the obstacle after every moment update."""

    def __init__(self, obstacle_mask=None, **kwargs):
        assert (obstacle_mask is not None)
        assert (np.sum(obstacle_mask) != 0)
        obstacle_mask = np.asfortranarray(obstacle_mask)
        self.obstacle_mask_host = obstacle_mask.astype(np.int32)
        self._zero_vel = True
        super(Pipe_Flow_Obstacles, self).__init__(**kwargs)

    def init_hydro(self):
        super(Pipe_Flow_Obstacles, self).init_hydro()
        if self.obstacle_mask_host.shape != (self.nx, self.ny):
            raise ValueError(f"obstacle_mask must have shape (nx, ny) = {(self.nx, self.ny)}")
        self.sim.set_mask(np.asarray(self.obstacle_mask_host).T)
        self.sim.zero_velocity_in_obstacle()
