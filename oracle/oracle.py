"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/liboracle_d2q9.so.

Arrays use the device layout of the reference, f[9][ny][nx] with x fastest
(D2Q9.cl:24-25); `to_ref_layout` / `from_ref_layout` convert to and from the
host layouts the reference classes expose:
  opencl_dim:  (nx, ny, 9) Fortran-order   (opencl_dim.py:165)
  cython_dim:  (9, nx, ny) C-order         (cython_dim.pyx:101)
"""
import ctypes as ct
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle_d2q9.so")

# opencl_dim.py:26-30 -- computed the way the reference computes them
cs = 1. / np.sqrt(3)
cs2 = cs ** 2
cs22 = 2 * cs2
two_cs4 = 2 * cs ** 4
cssq = 2.0 / 9.0

CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1])
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1])
W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)

BC_PIPE, BC_PERIODIC, BC_VELOCITY_YPERIODIC = 0, 1, 2


def build(force=False):
    src = [os.path.join(HERE, n) for n in ("d2q9_oracle.c", "d2q9_oracle_impl.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-o", LIB_PATH, src[0], "-lm"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ct.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ct.c_void_p) if a is not None else None


class CyParams(ct.Structure):
    _fields_ = [("nx", ct.c_int), ("ny", ct.c_int), ("omega", ct.c_double),
                ("inlet_rho", ct.c_double), ("outlet_rho", ct.c_double),
                ("cs2", ct.c_double), ("cs22", ct.c_double), ("cssq", ct.c_double),
                ("velocity_inlet", ct.c_int), ("u_w", ct.c_double), ("u_e", ct.c_double), ("old_api", ct.c_int)]


def feq_of(rho, u, v, dtype, incompressible=False):
    """Equilibrium from moments, scheme 'opencl' (D2Q9.cl:2-64 / D2Q9i.cl:2-63).  rho,u,v: (ny,nx)."""
    dtype = np.dtype(dtype)
    ny, nx = rho.shape
    r, uu, vv = (np.ascontiguousarray(a, dtype=dtype) for a in (rho, u, v))
    feq = np.empty((9, ny, nx), dtype=dtype)
    sfx = "f32" if dtype == np.float32 else "f64"
    if incompressible:
        getattr(lib(), "oracle_update_feq_i_" + sfx)(ct.c_int(nx), ct.c_int(ny), _p(r), _p(uu), _p(vv), _p(feq))
        return feq
    fn = getattr(lib(), "oracle_update_feq_" + sfx)
    fn(ct.c_int(nx), ct.c_int(ny), _p(r), _p(uu), _p(vv), _p(feq),
       ct.c_double(cs2), ct.c_double(cs22), ct.c_double(two_cs4))
    return feq


class OpenCLSchemeOracle:
    """State + stage methods of opencl_dim.Pipe_Flow[_Cylinder] on the CPU.

    f: (9, ny, nx) array (copied).  mask: (ny, nx) of {0,1} or None.

    bc=BC_VELOCITY_YPERIODIC selects LB_D2Q9/OLD/opencl.py's Pipe_Flow[_Obstacles]_PeriodicBC_VelocityInlet:
    D2Q9.cl's *_PeriodicBC_VelocityInlet kernels in the OLD step order (boundary pass BEFORE streaming);
    u0, v0 are the initial velocity arrays (parts of them are never rewritten, D2Q9.cl:357-371) and with
    a mask u, v are zeroed in the obstacle after every moment update (OLD/opencl.py:359-363).
    """

    def __init__(self, f0, omega, inlet_rho=1.0, outlet_rho=1.0, mask=None, bc=BC_PIPE,
                 dtype=np.float32, zero_obstacle_velocity=False, incompressible=False,
                 u_w=0.0, u_e=0.0, u0=None, v0=None):
        self.incompressible = bool(incompressible)      # D2Q9i.cl instead of D2Q9.cl
        self.dtype = np.dtype(dtype)
        self.sfx = "_f32" if self.dtype == np.float32 else "_f64"
        self.f = np.array(f0, dtype=self.dtype, order="C", copy=True)
        assert self.f.ndim == 3 and self.f.shape[0] == 9
        _, self.ny, self.nx = self.f.shape
        self.f_streamed = self.f.copy()          # opencl_dim.py:324-327
        self.mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.int32)
        self.bc = bc
        self.omega, self.inlet_rho, self.outlet_rho = float(omega), float(inlet_rho), float(outlet_rho)
        self.zero_obstacle_velocity = bool(zero_obstacle_velocity)
        shape2 = (self.ny, self.nx)
        self.rho = np.zeros(shape2, self.dtype)
        self.u = np.zeros(shape2, self.dtype)
        self.v = np.zeros(shape2, self.dtype)
        self.feq = np.zeros_like(self.f)
        self.u_w, self.u_e = float(u_w), float(u_e)
        if bc == BC_VELOCITY_YPERIODIC:
            assert not incompressible
            self.zero_obstacle_velocity = self.mask is not None
            if u0 is not None:
                self.u[...] = u0
            if v0 is not None:
                self.v[...] = v0

    def _fn(self, name):
        return getattr(lib(), name + self.sfx)

    def _dims(self):
        return ct.c_int(self.nx), ct.c_int(self.ny)

    def move(self):
        name = "oracle_move_periodic" if self.bc == BC_PERIODIC else "oracle_move"
        self._fn(name)(*self._dims(), _p(self.f), _p(self.f_streamed))

    def move_bcs(self):
        if self.bc == BC_VELOCITY_YPERIODIC:
            self._fn("oracle_move_bcs_vin")(*self._dims(), _p(self.f), ct.c_double(self.u_w), ct.c_double(self.u_e))
        if self.bc == BC_PIPE:
            self._fn("oracle_move_bcs_i" if self.incompressible else "oracle_move_bcs")(*self._dims(), _p(self.f), ct.c_double(self.inlet_rho),
                                        ct.c_double(self.outlet_rho))
        if self.mask is not None:
            self._fn("oracle_bounceback")(*self._dims(), _p(self.mask), _p(self.f))

    def update_hydro(self):
        if self.bc == BC_VELOCITY_YPERIODIC:
            self._fn("oracle_update_hydro_vin")(*self._dims(), _p(self.f), _p(self.rho), _p(self.u), _p(self.v),
                                                ct.c_double(self.u_w), ct.c_double(self.u_e))
        else:
            self._fn("oracle_update_hydro_i" if self.incompressible else "oracle_update_hydro")(
                *self._dims(), _p(self.f), _p(self.rho), _p(self.u), _p(self.v))
        if self.mask is not None and self.zero_obstacle_velocity:
            self._fn("oracle_zero_velocity")(*self._dims(), _p(self.mask), _p(self.u), _p(self.v))

    def update_feq(self):
        if self.incompressible:
            self._fn("oracle_update_feq_i")(*self._dims(), _p(self.rho), _p(self.u), _p(self.v), _p(self.feq))
            return
        self._fn("oracle_update_feq")(*self._dims(), _p(self.rho), _p(self.u), _p(self.v), _p(self.feq),
                                      ct.c_double(cs2), ct.c_double(cs22), ct.c_double(two_cs4))

    def collide_particles(self):
        self._fn("oracle_collide")(*self._dims(), _p(self.f), _p(self.feq), ct.c_double(self.omega))

    def run(self, n):
        if self.bc == BC_VELOCITY_YPERIODIC:
            self._fn("oracle_run_oldcl_vin")(*self._dims(), ct.c_int(int(n)), _p(self.f), _p(self.f_streamed),
                                             _p(self.mask), _p(self.rho), _p(self.u), _p(self.v), _p(self.feq),
                                             ct.c_double(self.omega), ct.c_double(self.u_w), ct.c_double(self.u_e),
                                             ct.c_double(cs2), ct.c_double(cs22), ct.c_double(two_cs4))
            return
        self._fn("oracle_run")(*self._dims(), ct.c_int(self.bc), ct.c_int(int(n)), _p(self.f),
                               _p(self.f_streamed), _p(self.mask), _p(self.rho), _p(self.u), _p(self.v),
                               _p(self.feq), ct.c_double(self.omega), ct.c_double(self.inlet_rho),
                               ct.c_double(self.outlet_rho), ct.c_double(cs2), ct.c_double(cs22),
                               ct.c_double(two_cs4), ct.c_int(int(self.zero_obstacle_velocity)),
                               ct.c_int(int(self.incompressible)))


class CythonSchemeOracle:
    """State + step of cython_dim.Pipe_Flow[_Cylinder] (cython_dim.pyx:204-359, :459-513).

    f: (9, ny, nx) float32; u, v: (ny, nx) float64 (the lagged velocity the
    first move_bcs reads); mask: (ny, nx) bool or None.  old_api=True selects the
    LB_D2Q9/OLD/cython.pyx flavour (see cy_params.old_api in d2q9_oracle.c); velocity_inlet=(u_w, u_e)
    additionally selects OLD's Pipe_Flow_PeriodicBC_VelocityInlet boundary family.
    """

    def __init__(self, f0, u0, v0, omega, inlet_rho, outlet_rho, mask=None, old_api=False, velocity_inlet=None):
        self.f = np.array(f0, dtype=np.float32, order="C", copy=True)
        _, self.ny, self.nx = self.f.shape
        self.u = np.array(u0, dtype=np.float64, order="C", copy=True)
        self.v = np.array(v0, dtype=np.float64, order="C", copy=True)
        self.rho = np.zeros((self.ny, self.nx), np.float32)
        self.feq = np.zeros_like(self.f)
        self.scratch = np.zeros_like(self.f)
        self.mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.p = CyParams(self.nx, self.ny, float(omega), float(inlet_rho), float(outlet_rho),
                          float(cs2), float(cs22), float(cssq),
                          int(velocity_inlet is not None), float((velocity_inlet or (0, 0))[0]),
                          float((velocity_inlet or (0, 0))[1]), int(bool(old_api)))

    def run(self, n):
        lib().oracle_cy_run(ct.byref(self.p), ct.c_int(int(n)), _p(self.f), _p(self.feq), _p(self.rho),
                            _p(self.u), _p(self.v), _p(self.mask), _p(self.scratch))


# ---- layout helpers -------------------------------------------------------
def from_opencl_host(a):
    """(nx, ny[, 9]) Fortran-order host array of opencl_dim -> ([9,] ny, nx) C-order."""
    a = np.asarray(a)
    return np.ascontiguousarray(a.transpose(2, 1, 0) if a.ndim == 3 else a.T)


def to_opencl_host(a):
    """([9,] ny, nx) -> (nx, ny[, 9]) Fortran-order, as opencl_dim.get_fields returns."""
    a = np.asarray(a)
    return np.asfortranarray(a.transpose(2, 1, 0) if a.ndim == 3 else a.T)


def from_cython_host(a):
    """(9, nx, ny) or (nx, ny) C-order arrays of cython_dim -> ([9,] ny, nx)."""
    a = np.asarray(a)
    return np.ascontiguousarray(a.transpose(0, 2, 1) if a.ndim == 3 else a.T)
