#!/usr/bin/env python
"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `python oracle/build_ref.py`):

    python tests/golden/make_golden.py

Every array is produced by the reference's own code from a seeded legacy NumPy RNG; nothing in
here comes from this repository's oracle or CUDA code:

  * cython_* / old_* : the compiled Cython classes (LB_D2Q9/dimensionless/cython_dim.pyx,
    LB_D2Q9/OLD/cython.pyx); host layout (9,nx,ny) C-order for f, (nx,ny) for rho/u/v;
  * opencl_* / oldcl_* : the OpenCL host classes (LB_D2Q9/dimensionless/opencl_dim.py,
    opencl_dim_D2Q9i.py, OLD/opencl.py) driving the reference's own kernel files D2Q9.cl / D2Q9i.cl,
    compiled as C and executed on the CPU by oracle/clshim (see oracle/shims/pyopencl); host layout
    (nx,ny,9) / (nx,ny) Fortran-order, as get_fields() returns them.

    python tests/golden/make_golden.py [cython|opencl|mask ...]     (default: everything)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, refload  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def snapshot(sim, tag, d):
    d[f"f_{tag}"] = np.array(sim.f, copy=True)
    d[f"rho_{tag}"] = np.array(sim.rho, copy=True)
    d[f"u_{tag}"] = np.array(sim.u, copy=True)
    d[f"v_{tag}"] = np.array(sim.v, copy=True)


def record(sim, steps, extra=None):
    d = dict(nx=sim.nx, ny=sim.ny, omega=float(sim.omega), inlet_rho=float(sim.inlet_rho),
             outlet_rho=float(sim.outlet_rho), steps=np.array(steps))
    if extra:
        d.update(extra)
    snapshot(sim, 0, d)
    done = 0
    for s in steps:
        sim.run(s - done)
        done = s
        snapshot(sim, s, d)
    return d


def cl_snapshot(sim, tag, d, with_feq=True):
    g = sim.get_fields() if hasattr(sim, "get_fields") else sim.get_fields_on_cpu()
    for k in ("f", "feq", "rho", "u", "v"):
        if k != "feq" or with_feq:
            d[f"{k}_{tag}"] = g[k]


def cl_record(sim, steps, extra=None):
    d = dict(nx=sim.nx, ny=sim.ny, omega=float(sim.omega), inlet_rho=float(sim.inlet_rho),
             outlet_rho=float(sim.outlet_rho), steps=np.array(steps))
    if extra:
        d.update(extra)
    cl_snapshot(sim, 0, d)
    done = 0
    for s in steps:
        sim.run(s - done)
        done = s
        cl_snapshot(sim, s, d, with_feq=(s == steps[-1]))     # feq only once: keeps the fixtures small
    return d


def opencl_goldens():
    """The reference's OpenCL path, end to end (host classes + kernel files), on the CPU emulation."""
    build_ref.build_cl()
    ocl = refload.opencl_dim()
    ocli = refload.opencl_dim_D2Q9i()
    oldcl = refload.old_opencl()
    ls = dict(two_d_local_size=(8, 8), three_d_local_size=(8, 8, 1))      # results do not depend on it

    # 5. opencl_dim.Pipe_Flow, 65x33 (same set-up as golden 1)
    kw = dict(diameter=1., rho=1., viscosity=0.0757, pressure_grad=-1., pipe_length=64 / 32., N=32,
              time_prefactor=6.)
    np.random.seed(10)
    with refload.quiet() as out:
        sim = ocl.Pipe_Flow(**kw, **ls)
    d = cl_record(sim, [1, 10, 100], dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=10))
    np.savez_compressed(os.path.join(OUT, "opencl_pipe_65x33.npz"), **d)

    # 6. opencl_dim.Pipe_Flow_Cylinder, 121x41: omega ~ 1.5, inlet rho ~ 1.06 (L = cylinder radius)
    kw = dict(cylinder_center=[0.75, 0.5], cylinder_radius=0.1, diameter=1., rho=1., viscosity=0.0556,
              pressure_grad=-10., pipe_length=3., N=4, time_prefactor=0.1)
    np.random.seed(11)
    with refload.quiet() as out:
        sim = ocl.Pipe_Flow_Cylinder(**kw, **ls)
    d = cl_record(sim, [1, 10, 100, 400],
                  dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=11, mask=np.array(sim.obstacle_mask_host)))
    np.savez_compressed(os.path.join(OUT, "opencl_cylinder_121x41.npz"), **d)

    # 7. opencl_dim_D2Q9i (D2Q9i.cl), pipe and cylinder.  As shipped this model overflows within a dozen
    #    steps at any parameters we tried (its equilibrium sums to rho**2); the vectors stop before that.
    kw = dict(diameter=1., rho=1., viscosity=.1, pressure_grad=-1., pipe_length=2., N=24, time_prefactor=1.)
    np.random.seed(12)
    with refload.quiet() as out:
        sim = ocli.Pipe_Flow(**kw, **ls)
    d = cl_record(sim, [1, 4, 8], dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=12))
    np.savez_compressed(os.path.join(OUT, "opencl_d2q9i_pipe_49x25.npz"), **d)
    kw = dict(cylinder_center=[0.75, 0.5], cylinder_radius=0.1, diameter=1., rho=1., viscosity=.2,
              pressure_grad=-1., pipe_length=3., N=4, time_prefactor=1.)
    np.random.seed(13)
    with refload.quiet() as out:
        sim = ocli.Pipe_Flow_Cylinder(**kw, **ls)
    d = cl_record(sim, [1, 4, 8],
                  dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=13, mask=np.array(sim.obstacle_mask_host)))
    np.savez_compressed(os.path.join(OUT, "opencl_d2q9i_cylinder_121x41.npz"), **d)

    # 8. OLD/opencl.py: the velocity-inlet / y-periodic classes, the only callers of D2Q9.cl's
    #    *_PeriodicBC_VelocityInlet kernels (step order BC -> stream -> moments -> feq -> collide).
    #    One obstacle touches the periodic row y = 0 on purpose.
    lx, ly = 60, 30
    kw = dict(lx=lx, ly=ly, omega=1.3, deltaP=-0.0, u_w=0.05)
    np.random.seed(14)
    with refload.quiet():
        sim = oldcl.Pipe_Flow_PeriodicBC_VelocityInlet(**kw, **ls)
    d = cl_record(sim, [1, 10, 100], dict(ctor_kwargs=repr(kw), seed=14, u_w=float(sim.u_w), u_e=float(sim.u_e)))
    np.savez_compressed(os.path.join(OUT, "oldcl_velocity_inlet_61x31.npz"), **d)
    mask = np.zeros((lx + 1, ly + 1), dtype=bool)
    mask[15:20, 10:18] = True
    mask[40:44, 0:4] = True
    np.random.seed(15)
    with refload.quiet():
        sim = oldcl.Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet(obstacle_mask=mask, **kw, **ls)
    d = cl_record(sim, [1, 10, 100, 400], dict(ctor_kwargs=repr(kw), seed=15, u_w=float(sim.u_w), u_e=float(sim.u_e),
                                                mask=mask.astype(np.uint8)))
    np.savez_compressed(os.path.join(OUT, "oldcl_velocity_inlet_obstacles_61x31.npz"), **d)

    # 9. OLD/opencl.py pressure-driven Pipe_Flow: the same kernels in the OLD order.  Kept short: the
    #    class is unstable as shipped (NaN within ~250 steps at these parameters).
    kw = dict(lx=48, ly=24, omega=1.0, deltaP=-0.001)
    np.random.seed(16)
    with refload.quiet():
        sim = oldcl.Pipe_Flow(**kw, **ls)
    d = cl_record(sim, [1, 10, 40], dict(ctor_kwargs=repr(kw), seed=16))
    np.savez_compressed(os.path.join(OUT, "oldcl_pipe_49x25.npz"), **d)


def listing():
    for n in sorted(os.listdir(OUT)):
        if n.endswith(".npz"):
            print(n, os.path.getsize(os.path.join(OUT, n)) // 1024, "KiB")


def main():
    what = set(sys.argv[1:]) or {"cython", "opencl", "mask"}
    if "opencl" in what:
        opencl_goldens()
    if "mask" in what:
        mask_golden()
    if "cython" in what:
        cython_goldens()
    listing()


def cython_goldens():
    build_ref.build()
    cd = refload.cython_dim()
    old = refload.old_cython()

    # 1. cython_dim.Pipe_Flow, 65x33, omega ~ 1.0003 (SURVEY 8d "better-conditioned C1", scaled down)
    kw = dict(diameter=1., rho=1., viscosity=0.0757, pressure_grad=-1., pipe_length=64 / 32., N=32,
              time_prefactor=6.)
    np.random.seed(0)
    with refload.quiet() as out:
        sim = cd.Pipe_Flow(**kw)
    d = record(sim, [1, 10, 100], dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=0))
    np.savez_compressed(os.path.join(OUT, "cython_pipe_65x33.npz"), **d)

    # 2. cython_dim.Pipe_Flow_Cylinder, 121x41 (the authors' cylinder set-up at N=4)
    kw = dict(cylinder_center=[0.75, 0.5], cylinder_radius=0.1, diameter=1., rho=1., viscosity=1.,
              pressure_grad=-10., pipe_length=3., N=4)
    np.random.seed(1)
    with refload.quiet() as out:
        sim = cd.Pipe_Flow_Cylinder(**kw)
    d = record(sim, [1, 10, 100],
               dict(printout=out.getvalue(), ctor_kwargs=repr(kw), seed=1,
                    mask=np.array(sim.obstacle_mask, dtype=np.uint8)))
    np.savez_compressed(os.path.join(OUT, "cython_cylinder_121x41.npz"), **d)

    # 3. OLD/cython.Pipe_Flow_Obstacles, 49x25, two rectangular obstacles (no init noise in OLD)
    lx, ly = 48, 24
    mask = np.zeros((lx + 1, ly + 1), dtype=bool)
    mask[10:14, 8:15] = True
    mask[30:33, 3:9] = True
    kw = dict(lx=lx, ly=ly, omega=1.2, deltaP=-0.02)
    np.random.seed(2)
    sim = old.Pipe_Flow_Obstacles(obstacle_mask=mask, **kw)
    d = record(sim, [1, 10, 100], dict(ctor_kwargs=repr(kw), seed=2, mask=mask.astype(np.uint8)))
    np.savez_compressed(os.path.join(OUT, "old_obstacles_49x25.npz"), **d)

    # 3b. OLD/cython velocity-inlet / y-periodic family, with and without obstacles
    lx, ly = 60, 30
    kw = dict(lx=lx, ly=ly, omega=1.3, deltaP=-0.0, u_w=0.05)
    np.random.seed(3)
    sim = old.Pipe_Flow_PeriodicBC_VelocityInlet(**kw)
    d = record(sim, [1, 10, 100], dict(ctor_kwargs=repr(kw), seed=3, u_w=float(sim.u_w), u_e=float(sim.u_e)))
    np.savez_compressed(os.path.join(OUT, "old_velocity_inlet_61x31.npz"), **d)
    mask = np.zeros((lx + 1, ly + 1), dtype=bool)
    mask[15:20, 10:18] = True
    mask[40:44, 2:6] = True
    np.random.seed(4)
    sim = old.Pipe_Flow_Obstacles_PeriodicBC_VelocityInlet(obstacle_mask=mask, **kw)
    d = record(sim, [1, 10, 100], dict(ctor_kwargs=repr(kw), seed=4, u_w=float(sim.u_w), u_e=float(sim.u_e),
                                       mask=mask.astype(np.uint8)))
    np.savez_compressed(os.path.join(OUT, "old_velocity_inlet_obstacles_61x31.npz"), **d)



def mask_golden():
    # 4. docs/cs205_binary.tif (800x400 px, {0,255}): the obstacle of BASELINE config 2, as a bit-packed
    #    (x, y) boolean mask.  No code in the reference loads this file; white = solid.
    sys.path.insert(0, os.path.join(ROOT, "2d-lb_b200"))
    from lb_b200 import masks
    m = masks.from_image("/root/reference/docs/cs205_binary.tif")
    assert m.shape == (800, 400) and 0.05 < m.mean() < 0.15
    np.savez_compressed(os.path.join(OUT, "cs205_binary_mask.npz"), source="docs/cs205_binary.tif", **masks.pack(m))


if __name__ == "__main__":
    main()
