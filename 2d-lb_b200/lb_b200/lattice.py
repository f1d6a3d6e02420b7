"""Low-level lattice object: one C-ABI handle (= one slab on one GPU).

Arrays exchanged with this class use the DEVICE layout of the reference:
f[9][ny][nx] / rho[ny][nx], x fastest (D2Q9.cl:24-25).  The reference's host arrays are
Fortran-order (nx, ny, 9) / (nx, ny) (opencl_dim.py:165), which is the same memory: `a.T` of
such an array is a C-contiguous view in device layout, no copy.
"""
import ctypes as ct

import numpy as np

from . import native as N

# Lattice constants exactly as the reference computes them (opencl_dim.py:26-30); they are
# handed to the library as doubles so both sides use the very same bits.
cs = 1. / np.sqrt(3)
cs2 = cs ** 2
cs22 = 2 * cs2
two_cs4 = 2 * cs ** 4

_BC = {"pipe": N.BC_PIPE, "periodic": N.BC_PERIODIC, "velocity_yperiodic": N.BC_VELOCITY_YPERIODIC}
_MATH = {"strict": N.MATH_STRICT, "fast": N.MATH_FAST}
_EDGE = {"boundary": N.EDGE_BOUNDARY, "wrap": N.EDGE_WRAP, "halo": N.EDGE_HALO}
_SCHEME = {"opencl": N.SCHEME_OPENCL, "cython": N.SCHEME_CYTHON, "cython_old": N.SCHEME_CYTHON_OLD,
           "opencl_old": N.SCHEME_OPENCL_OLD}
_FIELD = {"f": N.FIELD_F, "feq": N.FIELD_FEQ, "rho": N.FIELD_RHO, "u": N.FIELD_U, "v": N.FIELD_V}


def _ptr(a):
    return ct.c_void_p(a.ctypes.data)


def make_config(nx, ny, omega, inlet_rho=1.0, outlet_rho=1.0, bc="pipe", dtype=np.float32, math="strict", device=0,
                zero_obstacle_velocity=False, global_nx=None, x_offset=0, west_edge=None, east_edge=None, stream=None,
                scheme="opencl", model="d2q9", u_west=0.0, u_east=0.0):
    """The lb_config of include/lb_d2q9.h for one slab (or, for lb_multi_create, for the whole lattice)."""
    dtype = np.dtype(dtype)
    if dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    default_edge = "wrap" if bc == "periodic" else "boundary"
    cfg = N.LBConfig()
    cfg.struct_size = ct.sizeof(N.LBConfig)
    cfg.device = int(device)
    cfg.nx, cfg.ny = int(nx), int(ny)
    cfg.dtype = N.F32 if dtype == np.float32 else N.F64
    cfg.bc = _BC[bc]
    cfg.math = _MATH[math]
    cfg.zero_obstacle_velocity = int(bool(zero_obstacle_velocity))
    cfg.global_nx = int(global_nx if global_nx is not None else nx)
    cfg.x_offset = int(x_offset)
    cfg.west_edge = _EDGE[west_edge or default_edge]
    cfg.east_edge = _EDGE[east_edge or default_edge]
    cfg.scheme = _SCHEME[scheme]
    cfg.model = {"d2q9": N.MODEL_D2Q9, "d2q9i": N.MODEL_D2Q9I}[model]
    cfg.omega, cfg.inlet_rho, cfg.outlet_rho = float(omega), float(inlet_rho), float(outlet_rho)
    cfg.cs2, cfg.cs22, cfg.two_cs4 = float(cs2), float(cs22), float(two_cs4)
    cfg.u_west, cfg.u_east = float(u_west), float(u_east)
    cfg.stream = ct.c_void_p(stream) if stream else None
    return cfg


class Lattice:
    """D2Q9 BGK lattice on one CUDA device.

    Parameters mirror what the reference's kernels take (omega, inlet/outlet density,
    obstacle mask, populations) rather than the physical parameters of the simulation
    classes; `lb_b200.dimensionless` builds on this.

    bc     'pipe' (pressure inlet/outlet + walls, D2Q9.cl:173-261), 'periodic', or 'velocity_yperiodic'
           (imposed inlet/outlet velocity u_west/u_east, rows 0 and ny-1 exchanged; schemes 'cython_old'
           and 'opencl_old')
    math   'strict' (default: D2Q9.cl's arithmetic operation for operation, bit-identical to the CPU
           oracle, and HBM-bound like 'fast') or 'fast' (FMA + reciprocal constants, ~30% fewer
           instructions, agrees to rounding)
    dtype  np.float32 (the reference's precision) or np.float64
    model  'd2q9' (D2Q9.cl) or 'd2q9i' (the incompressible variant D2Q9i.cl; scheme 'opencl', pipe flow)
    scheme 'opencl' (default: the step order of opencl_dim.py), 'cython' or
           'cython_old' (the reference's CPU classes, cython_dim.pyx / OLD/cython.pyx, including
           their mixed-precision arithmetic; float32 populations, float64 u and v), or 'opencl_old'
           (D2Q9.cl's velocity-inlet kernels in the step order of OLD/opencl.py; float32 throughout)
    """

    def __init__(self, nx, ny, omega, inlet_rho=1.0, outlet_rho=1.0, mask=None, f0=None, bc="pipe",
                 dtype=np.float32, math="strict", device=0, zero_obstacle_velocity=False,
                 global_nx=None, x_offset=0, west_edge=None, east_edge=None, stream=None, scheme="opencl",
                 model="d2q9", u_west=0.0, u_east=0.0):
        self._h = None
        self.scheme = scheme
        self.nx, self.ny = int(nx), int(ny)
        self.dtype = np.dtype(dtype)
        self.bc, self.math = bc, math
        cfg = make_config(nx, ny, omega, inlet_rho, outlet_rho, bc=bc, dtype=dtype, math=math, device=device,
                          zero_obstacle_velocity=zero_obstacle_velocity, global_nx=global_nx, x_offset=x_offset,
                          west_edge=west_edge, east_edge=east_edge, stream=stream, scheme=scheme, model=model,
                          u_west=u_west, u_east=u_east)
        self.model = model
        self.cfg = cfg
        self.omega, self.inlet_rho, self.outlet_rho = cfg.omega, cfg.inlet_rho, cfg.outlet_rho
        self.device = cfg.device
        self._stream_owner = None     # set by whoever lends this lattice a stream: keeps the lender alive
        h = ct.c_void_p()
        N.check(N.lib().lb_create(ct.byref(cfg), ct.byref(h)))
        self._h = h
        if mask is not None:
            self.set_mask(mask)
        if f0 is not None:
            self.upload_f(f0)

    from_lattice = classmethod(lambda cls, *a, **k: cls(*a, **k))

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            N.lib().lb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _call(self, fn, *args):
        if self._h is None:
            raise N.LBError(-3, "lattice is closed")
        N.check(getattr(N.lib(), fn)(self._h, *args), self._h)

    # -- uploads ----------------------------------------------------------------------
    def set_mask(self, mask):
        """mask: (ny, nx); entries equal to 1 are solid (D2Q9.cl:410).  None removes it."""
        if mask is None:
            self._call("lb_set_mask", None, 1)
            return
        m = np.asarray(mask)
        if m.shape != (self.ny, self.nx):
            raise ValueError(f"mask must have shape (ny, nx) = {(self.ny, self.nx)}, got {m.shape}")
        m = np.ascontiguousarray(m == 1, dtype=np.uint8)     # solid <=> exactly 1 (D2Q9.cl:410), whatever the dtype
        self._call("lb_set_mask", _ptr(m), 1)

    def upload_f(self, f):
        """f: (9, ny, nx) populations; seeds both ping-pong buffers (opencl_dim.py:324-327)."""
        a = np.ascontiguousarray(f, dtype=self.dtype)
        if a.shape != (9, self.ny, self.nx):
            raise ValueError(f"f must have shape (9, ny, nx) = {(9, self.ny, self.nx)}, got {a.shape}")
        self._call("lb_upload_f", _ptr(a))

    def upload_moments(self, rho=None, u=None, v=None):
        arrs = []
        for k, a in enumerate((rho, u, v)):
            if a is None:
                arrs.append(None)
                continue
            b = np.ascontiguousarray(a, dtype=self.dtype if k == 0 else self.uv_dtype)
            if b.shape != (self.ny, self.nx):
                raise ValueError("moment fields must have shape (ny, nx)")
            arrs.append(b)
        self._call("lb_upload_moments", *[(_ptr(a) if a is not None else None) for a in arrs])

    @property
    def uv_dtype(self):
        """dtype of the u / v fields: float64 for the cython schemes (as in the reference)."""
        return np.dtype(np.float64) if self.scheme in ("cython", "cython_old") else self.dtype

    def field_dtype(self, field):
        return self.uv_dtype if field in ("u", "v") else self.dtype

    # -- the hot path -----------------------------------------------------------------
    def run(self, n, sync=True):
        """n fused collide-and-stream steps (opencl_dim.Pipe_Flow.run); one sync at the end."""
        self._call("lb_step", int(n))
        if sync:
            self.sync()

    def sync(self):
        self._call("lb_sync")

    def run_streamed(self, f, n, rho=None, u=None, v=None):
        """upload_f(f) + run(n) + download of rho, u, v into the given arrays (None = not wanted), pipelined
        by row bands so that the host-to-device copy, the n steps and the device-to-host copies overlap
        (lb_run_streamed).  Pass page-locked arrays; same bits as the three separate calls."""
        a = np.ascontiguousarray(f, dtype=self.dtype)
        if a.shape != (9, self.ny, self.nx):
            raise ValueError(f"f must have shape (9, ny, nx) = {(9, self.ny, self.nx)}, got {a.shape}")
        outs = []
        for name, o in (("rho", rho), ("u", u), ("v", v)):
            if o is not None and (o.shape != (self.ny, self.nx) or o.dtype != self.field_dtype(name) or not o.flags.c_contiguous):
                raise ValueError(f"{name} must be a C-contiguous (ny, nx) array of dtype {self.field_dtype(name)}")
            outs.append(_ptr(o) if o is not None else None)
        self._call("lb_run_streamed", _ptr(a), int(n), *outs)

    # -- readback ---------------------------------------------------------------------
    def download(self, field, out=None):
        """Device layout array of `field` in {'f','feq','rho','u','v'}."""
        shape = (9, self.ny, self.nx) if field in ("f", "feq") else (self.ny, self.nx)
        dt = self.field_dtype(field)
        if out is None:
            out = np.empty(shape, dtype=dt)
        else:
            if out.shape != shape or out.dtype != dt or not out.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous array of the field's shape and dtype")
        self._call("lb_download", _FIELD[field], _ptr(out))
        return out

    def download_strided(self, field, stride_x, stride_y=None):
        """Every stride_x-th column / stride_y-th row of rho, u or v (gathered on the device): frames for
        long visual runs without a full-field copy per frame."""
        stride_y = stride_x if stride_y is None else stride_y
        if field not in ("rho", "u", "v"):
            raise ValueError("download_strided serves rho, u and v")
        out = np.empty(((self.ny + stride_y - 1) // stride_y, (self.nx + stride_x - 1) // stride_x),
                       dtype=self.field_dtype(field))
        self._call("lb_download_strided", _FIELD[field], int(stride_x), int(stride_y), _ptr(out))
        return out

    def fields(self):
        return {k: self.download(k) for k in ("f", "feq", "rho", "u", "v")}

    # -- single stages (D2Q9.cl kernel by kernel) --------------------------------------
    def move(self):
        self._call("lb_stage_move")

    def move_bcs(self):
        self._call("lb_stage_move_bcs")

    def update_hydro(self):
        self._call("lb_stage_update_hydro")

    def update_feq(self):
        self._call("lb_stage_update_feq")

    def collide_particles(self):
        self._call("lb_stage_collide")

    def zero_velocity_in_obstacle(self):
        self._call("lb_stage_zero_velocity")

    # -- synthetic initialisers / diagnostics -------------------------------------------
    def init_synthetic(self, kind="pipe_ramp", u0=0.0, amplitude=0.0, seed=0):
        k = {"pipe_ramp": N.SYNTH_PIPE_RAMP, "shear_layers": N.SYNTH_SHEAR_LAYERS}[kind]
        self._call("lb_init_synthetic", k, float(u0), float(amplitude), ct.c_uint64(int(seed)))

    def set_mask_disk(self, cx, cy, r):
        self._call("lb_set_mask_disk", float(cx), float(cy), float(r))

    def total_mass(self):
        out = ct.c_double()
        self._call("lb_total_mass", ct.byref(out))
        return out.value

    def set_temporal_blocking(self, shape):
        """Two lattice updates per pass through HBM (csrc/lb_march.cuh): `shape` is 'auto' / -1 (default:
        the measured-best shape on large lattices), 'off' / 0, a shape index or a name such as
        'march.w4b4.s64'.  Bit-identical results; 'opencl' scheme, single slabs and halo-connected slabs."""
        if isinstance(shape, str):
            if shape == "auto":
                shape = -1
            else:
                names = [N.lib().lb_tb2_shape_name(k).decode() for k in range(N.lib().lb_tb2_shape_count())]
                shape = names.index(shape)
        self._call("lb_set_temporal_blocking", int(shape))

    @property
    def temporal_blocking(self):
        """Name of the two-step tile run() uses on this lattice ('off' = the one-step kernel)."""
        k = N.lib().lb_temporal_blocking(self._h)
        return N.lib().lb_tb2_shape_name(k).decode()

    @property
    def segment_rows(self):
        """Rows per (strip, segment) work item of the marching kernel's launches on this lattice (0: one-update kernel)."""
        return int(N.lib().lb_segment_rows(self._h))

    def copy_ceiling_ms(self, reps=10):
        """ms per launch of an arithmetic-free kernel with the fused step's memory access pattern (the
        practical HBM ceiling of this device for this lattice; populations are left untouched)."""
        out = ct.c_double()
        self._call("lb_selftest_copy", int(reps), ct.byref(out))
        return out.value

    def checksum(self):
        """Order-independent exact checksum of the populations (64-bit sum of their bit patterns)."""
        out = ct.c_uint64()
        self._call("lb_checksum", ct.byref(out))
        return out.value

    @property
    def launch_count(self):
        return int(N.lib().lb_launch_count(self._h))

    def set_variant(self, variant):
        """variant: index or name from `native.variants()`; -1 restores the default."""
        if isinstance(variant, str):
            variant = N.variants().index(variant)
        self._call("lb_set_variant", int(variant))

    @property
    def stream_ptr(self):
        """cudaStream_t (as int) this lattice enqueues on."""
        return N.lib().lb_stream(self._h)

    def device_ptr(self, field):
        p, pitch = ct.c_void_p(), ct.c_int64()
        self._call("lb_device_ptr", _FIELD[field], ct.byref(p), ct.byref(pitch))
        return p.value, pitch.value

    # -- halo ---------------------------------------------------------------------------
    def halo_ipc_handle(self):
        buf = ct.create_string_buffer(N.IPC_HANDLE_BYTES)
        self._call("lb_halo_ipc_handle", buf)
        return buf.raw

    def halo_connect_ipc(self, side, handle_bytes, peer_device=0):
        buf = ct.create_string_buffer(handle_bytes, N.IPC_HANDLE_BYTES)
        self._call("lb_halo_connect_ipc", {"west": N.WEST, "east": N.EAST}[side], buf, int(peer_device))

    def halo_connect_local(self, side, peer):
        self._call("lb_halo_connect_local", {"west": N.WEST, "east": N.EAST}[side], peer._h)

    def halo_prime(self):
        self._call("lb_halo_prime")

    def set_halo_timeout(self, seconds):
        self._call("lb_set_halo_timeout", float(seconds))


def split_slabs(global_nx, parts):
    """Column ranges [(x_offset, nx), ...] of an x-slab decomposition into `parts` slabs.

    Remainder columns go to the first slabs, so widths differ by at most one.
    """
    base, rem = divmod(int(global_nx), int(parts))
    if base < 2:
        raise ValueError("each slab needs at least 2 columns")
    out, x = [], 0
    for r in range(parts):
        w = base + (1 if r < rem else 0)
        out.append((x, w))
        x += w
    return out


def slab_edges(rank, parts, bc):
    """(west_edge, east_edge) names of slab `rank` of `parts` for boundary family `bc`."""
    if parts == 1:
        e = "wrap" if bc == "periodic" else "boundary"
        return e, e
    if bc == "periodic":
        return "halo", "halo"
    return ("boundary" if rank == 0 else "halo"), ("boundary" if rank == parts - 1 else "halo")


class _BorrowedSlab(Lattice):
    """One slab of an lb_multi handle, seen through the single-slab interface (tuning / diagnostics).  The
    multi handle owns it: closing this view frees nothing."""

    def __init__(self, handle, cfg_like, nx, ny, dtype, scheme):
        self._h = handle
        self.nx, self.ny, self.dtype, self.scheme = nx, ny, np.dtype(dtype), scheme
        self.cfg = cfg_like
        self._stream_owner = None

    def close(self):
        self._h = None


class LocalSlabs:
    """`parts` x-slabs of one lattice in ONE process, with the same interface as `Lattice` (device-layout
    GLOBAL arrays in, global arrays out).  A thin caller of the C library's multi-device handle (lb_multi_*,
    include/lb_d2q9.h), which owns the slabs, their streams and peer mappings, keeps the ghost columns primed
    and runs the step loop:

    * one slab per DEVICE (`devices=[0, 1, ...]`): the slabs advance asynchronously and synchronise among
      themselves through the peer-memory flags -- the single-process multi-GPU path behind
      `Pipe_Flow(..., devices=[...])`;
    * several slabs on ONE device ("virtual ranks"): they share a stream and advance in lock-step, one launch
      at a time.  Used to prove that the decomposition is arithmetic-neutral.
    """

    def __init__(self, global_nx, ny, parts=None, devices=None, **kw):
        self._m = None
        if parts is None:
            parts = len(devices)
        self.global_nx, self.ny, self.parts = int(global_nx), int(ny), int(parts)
        self.nx = self.global_nx
        self.bc = kw.get("bc", "pipe")
        self.scheme = kw.get("scheme", "opencl")
        devices = list(devices) if devices else [kw.pop("device", 0)] * parts
        kw.pop("device", None)
        if len(devices) != parts:
            raise ValueError("one device per slab")
        self.devices = devices
        self.ranges = split_slabs(global_nx, parts)
        cfg = make_config(global_nx, ny, **kw)
        self.cfg = cfg
        self.dtype = np.dtype(kw.get("dtype", np.float32))
        self.uv_dtype = np.dtype(np.float64) if self.scheme in ("cython", "cython_old") else self.dtype
        h = ct.c_void_p()
        N.check_multi(N.lib().lb_multi_create(ct.byref(cfg), self.parts, (ct.c_int * self.parts)(*devices), ct.byref(h)))
        self._m = h
        self.slabs = []
        for k in range(self.parts):
            sh, x0, w = ct.c_void_p(), ct.c_int(), ct.c_int()
            self._call("lb_multi_slab", k, ct.byref(sh), ct.byref(x0), ct.byref(w))
            assert (x0.value, w.value) == self.ranges[k]
            self.slabs.append(_BorrowedSlab(sh, cfg, w.value, self.ny, self.dtype, self.scheme))

    def _call(self, fn, *args):
        if self._m is None:
            raise N.LBError(-3, "lattice is closed")
        N.check_multi(getattr(N.lib(), fn)(self._m, *args), self._m)

    def close(self):
        if self._m is not None:
            N.lib().lb_multi_destroy(self._m)      # drains every slab before freeing any arena
            self._m = None
            for s in self.slabs:
                s.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def field_dtype(self, field):
        return self.uv_dtype if field in ("u", "v") else self.dtype

    # -- uploads (global device-layout arrays) -------------------------------------------------------
    def set_mask(self, mask):
        if mask is None:
            self._call("lb_multi_set_mask", None, 1)
            return
        m = np.asarray(mask)
        if m.shape != (self.ny, self.nx):
            raise ValueError(f"mask must have shape (ny, nx) = {(self.ny, self.nx)}, got {m.shape}")
        m = np.ascontiguousarray(m == 1, dtype=np.uint8)
        self._call("lb_multi_set_mask", _ptr(m), 1)

    def upload_f(self, f):
        a = np.ascontiguousarray(f, dtype=self.dtype)
        if a.shape != (9, self.ny, self.nx):
            raise ValueError(f"f must have shape (9, ny, nx) = {(9, self.ny, self.nx)}, got {a.shape}")
        self._call("lb_multi_upload_f", _ptr(a))

    def upload_moments(self, rho=None, u=None, v=None):
        arrs = []
        for k, a in enumerate((rho, u, v)):
            arrs.append(None if a is None else np.ascontiguousarray(a, dtype=self.dtype if k == 0 else self.uv_dtype))
            if arrs[-1] is not None and arrs[-1].shape != (self.ny, self.nx):
                raise ValueError("moment fields must have shape (ny, nx)")
        self._call("lb_multi_upload_moments", *[(_ptr(a) if a is not None else None) for a in arrs])

    def prime(self):
        """Publish every slab's boundary columns to its neighbours (only needed after changing a slab's state
        through `slabs[k]`; uploads and initialisers do it themselves)."""
        self._call("lb_multi_prime")

    # -- hot path ---------------------------------------------------------------------------------
    def run(self, n, sync=True):
        self._call("lb_multi_step", int(n))
        if sync:
            self.sync()

    def sync(self):
        self._call("lb_multi_sync")

    def set_temporal_blocking(self, shape):
        if isinstance(shape, str):
            if shape == "auto":
                shape = -1
            else:
                names = [N.lib().lb_tb2_shape_name(k).decode() for k in range(N.lib().lb_tb2_shape_count())]
                shape = names.index(shape)
        self._call("lb_multi_set_temporal_blocking", int(shape))

    @property
    def temporal_blocking(self):
        return N.lib().lb_tb2_shape_name(N.lib().lb_multi_temporal_blocking(self._m)).decode()

    # -- stages that need no halo -----------------------------------------------------------------
    def update_feq(self):
        for s in self.slabs:
            s.update_feq()

    def zero_velocity_in_obstacle(self):
        for s in self.slabs:
            s.zero_velocity_in_obstacle()

    def _no_single_stage(self, *_):
        raise N.LBError(-3, "single stages are not available on a slab-decomposed lattice; use run()")

    move = move_bcs = update_hydro = collide_particles = _no_single_stage

    # -- readback ------------------------------------------------------------------------------------
    def download(self, field, out=None):
        shape = (9, self.ny, self.nx) if field in ("f", "feq") else (self.ny, self.nx)
        dt = self.field_dtype(field)
        if out is None:
            out = np.empty(shape, dtype=dt)
        elif out.shape != shape or out.dtype != dt or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous array of the field's shape and dtype")
        self._call("lb_multi_download", _FIELD[field], _ptr(out))
        return out

    def fields(self):
        return {k: self.download(k) for k in ("f", "feq", "rho", "u", "v")}

    def total_mass(self):
        out = ct.c_double()
        self._call("lb_multi_total_mass", ct.byref(out))
        return out.value

    def checksum(self):
        out = ct.c_uint64()
        self._call("lb_multi_checksum", ct.byref(out))
        return out.value

    @property
    def launch_count(self):
        return int(N.lib().lb_multi_launch_count(self._m))

    def init_synthetic(self, kind="pipe_ramp", u0=0.0, amplitude=0.0, seed=0):
        k = {"pipe_ramp": N.SYNTH_PIPE_RAMP, "shear_layers": N.SYNTH_SHEAR_LAYERS}[kind]
        self._call("lb_multi_init_synthetic", k, float(u0), float(amplitude), ct.c_uint64(int(seed)))

    def set_mask_disk(self, cx, cy, r):
        self._call("lb_multi_set_mask_disk", float(cx), float(cy), float(r))
