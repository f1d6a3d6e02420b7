"""The oracle of the path being replaced, pinned against the reference's OWN OpenCL code (CPU, `-m "not gpu"`).

pyopencl and an OpenCL runtime do not exist in this image, so the reference's kernel files
(LB_D2Q9/D2Q9.cl, D2Q9i.cl) are compiled as C by gcc through oracle/clshim/opencl_c.h and executed
work-item by work-item on the host, driven by the reference's unmodified host classes through the
`pyopencl` stand-in in oracle/shims/ (tests/golden/make_golden.py, `opencl_*` / `oldcl_*` vectors).

1. committed golden vectors of opencl_dim.Pipe_Flow / Pipe_Flow_Cylinder (and the D2Q9i twins)
   == OpenCLSchemeOracle, BIT FOR BIT, at every recorded step;
2. every kernel of the compiled D2Q9.cl, launched directly on random inputs, == the matching oracle
   stage function, BIT FOR BIT (runs wherever oracle/_ref/clshim travelled, the GPU box included);
3. where the reference tree is mounted: the live host classes at another size and seed;
4. the emulation layer itself: NDRange ids, work-group barriers, argument checking.
"""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def _same(a, b):
    """Bit-identical, NaN payloads excepted (D2Q9i overflows; both sides must then be non-finite
    at the same places)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    nan = np.isnan(a)
    if not np.array_equal(nan, np.isnan(b)):
        return False
    return np.array_equal(_bits(a)[~nan], _bits(b)[~nan])


@pytest.fixture(scope="module")
def cl():
    shim = os.path.join(ROOT, "oracle", "shims")
    if shim not in sys.path:
        sys.path.append(shim)
    import pyopencl
    assert pyopencl.VERSION_TEXT.startswith("clshim"), "these tests are written for the CPU emulation"
    return pyopencl


# ------------------------------------------------------------------------------------------------
# 1. golden vectors of the reference's OpenCL path
@pytest.mark.parametrize("name,incompressible", [("opencl_pipe_65x33.npz", False), ("opencl_cylinder_121x41.npz", False),
                                                 ("opencl_d2q9i_pipe_49x25.npz", True),
                                                 ("opencl_d2q9i_cylinder_121x41.npz", True)])
def test_opencl_scheme_matches_reference_golden_bitexact(orc, name, incompressible):
    g = _load(name)
    mask = orc.from_opencl_host(g["mask"]) if "mask" in g.files else None
    # opencl_dim.py passes np.float32(omega), np.float32(inlet_rho), np.float32(outlet_rho) (:334, :369)
    o = orc.OpenCLSchemeOracle(orc.from_opencl_host(g["f_0"]), np.float32(g["omega"]), np.float32(g["inlet_rho"]),
                               np.float32(g["outlet_rho"]), mask=mask, incompressible=incompressible,
                               # opencl_dim_D2Q9i.py zeroes u, v in the obstacle after every update_hydro
                               zero_obstacle_velocity=incompressible and mask is not None)
    done = 0
    for s in g["steps"]:
        o.run(int(s) - done)
        done = int(s)
        for k in ("f", "rho", "u", "v"):
            assert _same(getattr(o, k), orc.from_opencl_host(g[f"{k}_{s}"])), f"{k} after {s} steps"
    assert _same(o.feq, orc.from_opencl_host(g[f"feq_{done}"]))
    if not incompressible:
        assert np.isfinite(o.f).all()


def test_golden_initial_state_is_the_equilibrium_times_noise(orc):
    """f_0 of the golden = float32(feq(rho ramp, 0, 0) * (1 + 1e-3 randn)) (opencl_dim.py:270-327): the
    oracle's equilibrium reproduces the device-computed feq_0 bit for bit."""
    g = _load("opencl_pipe_65x33.npz")
    rho0 = orc.from_opencl_host(g["rho_0"])
    feq = orc.feq_of(rho0, np.zeros_like(rho0), np.zeros_like(rho0), np.float32)
    assert _same(feq, orc.from_opencl_host(g["feq_0"]))
    nx = int(g["nx"])
    ramp = (np.float64(g["inlet_rho"]) - np.arange(nx) * (np.float64(g["inlet_rho"]) - 1.0) / float(nx)).astype(np.float32)
    assert np.array_equal(rho0, ramp[None, :].repeat(rho0.shape[0], 0))


def test_old_opencl_pipe_is_the_same_kernels_in_the_old_order(orc):
    """OLD/opencl.py runs move_bcs -> move -> update_hydro -> update_feq -> collide (:246-255) on the same
    D2Q9.cl: the oracle's stage functions in that order reproduce its golden vector bit for bit --
    including the populations `move` never writes, which keep the initial value for ever."""
    g = _load("oldcl_pipe_49x25.npz")
    o = orc.OpenCLSchemeOracle(orc.from_opencl_host(g["f_0"]), np.float32(g["omega"]), np.float32(g["inlet_rho"]),
                               np.float32(g["outlet_rho"]))
    done = 0
    for s in g["steps"]:
        for _ in range(int(s) - done):
            o.move_bcs(); o.move(); o.update_hydro(); o.update_feq(); o.collide_particles()
        done = int(s)
        for k in ("f", "rho", "u", "v"):
            assert _same(getattr(o, k), orc.from_opencl_host(g[f"{k}_{s}"])), f"{k} after {s} steps"


@pytest.mark.parametrize("name", ["oldcl_velocity_inlet_61x31.npz", "oldcl_velocity_inlet_obstacles_61x31.npz"])
def test_old_opencl_velocity_inlet_matches_reference_golden_bitexact(orc, name):
    """OLD/opencl.py's Pipe_Flow[_Obstacles]_PeriodicBC_VelocityInlet -- the only callers of D2Q9.cl:263-374 --
    against the oracle's restatement (bc=BC_VELOCITY_YPERIODIC).  One obstacle of the second vector
    touches the periodic row y = 0."""
    g = _load(name)
    mask = orc.from_opencl_host(g["mask"]).astype(np.int32) if "mask" in g.files else None
    o = orc.OpenCLSchemeOracle(orc.from_opencl_host(g["f_0"]), np.float32(g["omega"]), mask=mask,
                               bc=orc.BC_VELOCITY_YPERIODIC, u_w=np.float32(g["u_w"]), u_e=np.float32(g["u_e"]),
                               u0=orc.from_opencl_host(g["u_0"]), v0=orc.from_opencl_host(g["v_0"]))
    if mask is not None:
        assert mask[0].any()
    done = 0
    for s in g["steps"]:
        o.run(int(s) - done)
        done = int(s)
        for k in ("f", "rho", "u", "v"):
            assert _same(getattr(o, k), orc.from_opencl_host(g[f"{k}_{s}"])), f"{k} after {s} steps"
    assert _same(o.feq, orc.from_opencl_host(g[f"feq_{done}"]))
    assert np.isfinite(o.f).all() and float(np.abs(o.u).max()) < 0.2


def test_old_opencl_velocity_inlet_stages_equal_run(orc):
    from util import pipe_case
    f0, mask = pipe_case(orc, 45, 22, mask="touching", seed=5)
    kw = dict(mask=mask.astype(np.int32), bc=orc.BC_VELOCITY_YPERIODIC, u_w=np.float32(0.04), u_e=np.float32(0.04),
              u0=np.full((22, 45), np.float32(0.04)))
    a = orc.OpenCLSchemeOracle(f0, np.float32(1.1), **kw)
    b = orc.OpenCLSchemeOracle(f0, np.float32(1.1), **kw)
    a.run(7)
    for _ in range(7):
        b.move_bcs(); b.move(); b.update_hydro(); b.update_feq(); b.collide_particles()
    assert _same(a.f, b.f) and _same(a.u, b.u) and _same(a.rho, b.rho)


# ------------------------------------------------------------------------------------------------
# 2. the compiled kernels of D2Q9.cl, one by one, against the oracle's stage functions
@pytest.fixture(scope="module")
def d2q9(cl):
    try:
        return cl.Program.from_cache(cl.Context(), "D2Q9")
    except cl.Error as exc:
        pytest.skip(str(exc))


W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4, dtype=np.float32)
CX = np.array([0, 1, 0, -1, 0, 1, -1, -1, 1], dtype=np.int32)
CY = np.array([0, 0, 1, 0, -1, 1, 1, -1, -1], dtype=np.int32)


def _buf(cl, a):
    return cl.Buffer(None, cl.mem_flags.READ_WRITE | cl.mem_flags.COPY_HOST_PTR, hostbuf=np.ascontiguousarray(a))


def _read(cl, buf, shape, dtype=np.float32):
    out = np.empty(shape, dtype)
    cl.enqueue_copy(None, out, buf)
    return out


def _gsize(n, l):
    return tuple(-(-a // b) * b for a, b in zip(n, l))


@pytest.mark.parametrize("nx,ny,lsz", [(37, 19, (8, 4)), (64, 32, (32, 32)), (5, 3, (1, 1))])
def test_compiled_cl_kernels_equal_oracle_stages(orc, cl, d2q9, nx, ny, lsz):
    from util import pipe_case
    f0, mask = pipe_case(orc, nx, ny, mask="touching", seed=nx)
    mask = mask.astype(np.int32)
    omega, rin, rout = np.float32(1.37), np.float32(1.02), np.float32(0.99)
    o = orc.OpenCLSchemeOracle(f0, omega, rin, rout, mask=mask)
    l3 = lsz + (3,) if lsz != (32, 32) else lsz + (1,)
    g2, g3 = _gsize((nx, ny), lsz), _gsize((nx, ny, 9), l3)
    f, fs = _buf(cl, f0), _buf(cl, f0)
    u, v, rho = (_buf(cl, np.zeros((ny, nx), np.float32)) for _ in range(3))
    feq = _buf(cl, np.zeros_like(f0))
    m = _buf(cl, mask)
    w, cx, cy = _buf(cl, W), _buf(cl, CX), _buf(cl, CY)
    loc = [cl.LocalMemory(4 * lsz[0] * lsz[1]) for _ in range(3)]
    i32 = np.int32
    cs = 1. / np.sqrt(3)
    for step in range(3):
        d2q9.move(None, g3, l3, f, fs, cx, cy, i32(nx), i32(ny)).wait()
        d2q9.copy_buffer(None, g3, l3, fs, f, i32(nx), i32(ny)).wait()
        o.move()
        assert _same(_read(cl, f, f0.shape), o.f), f"move, step {step}"
        d2q9.move_bcs(None, g2, lsz, f, u, rin, rout, i32(nx), i32(ny)).wait()
        d2q9.bounceback_in_obstacle(None, g2, lsz, m, f, i32(nx), i32(ny)).wait()
        o.move_bcs()
        assert _same(_read(cl, f, f0.shape), o.f), f"move_bcs + bounceback, step {step}"
        d2q9.update_hydro(None, g2, lsz, f, u, v, rho, rin, rout, i32(nx), i32(ny)).wait()
        o.update_hydro()
        for b, want in ((rho, o.rho), (u, o.u), (v, o.v)):
            assert _same(_read(cl, b, (ny, nx)), want), f"update_hydro, step {step}"
        d2q9.update_feq(None, g3, l3, feq, u, v, rho, *loc, w, cx, cy, np.float32(cs), np.float32(cs ** 2),
                        np.float32(2 * cs ** 2), np.float32(2 * cs ** 4), i32(nx), i32(ny)).wait()
        o.update_feq()
        assert _same(_read(cl, feq, f0.shape), o.feq), f"update_feq, step {step}"
        d2q9.collide_particles(None, g3, l3, f, feq, omega, i32(nx), i32(ny)).wait()
        o.collide_particles()
        assert _same(_read(cl, f, f0.shape), o.f), f"collide_particles, step {step}"
    # the velocity-inlet pair (D2Q9.cl:263-374), on the state reached above
    uw, ue = np.float32(0.05), np.float32(0.03)
    ov = orc.OpenCLSchemeOracle(o.f, omega, bc=orc.BC_VELOCITY_YPERIODIC, u_w=uw, u_e=ue, u0=o.u, v0=o.v)
    if ny >= 4:
        d2q9.move_bcs_PeriodicBC_VelocityInlet(None, g2, lsz, f, u, uw, ue, i32(nx), i32(ny)).wait()
        ov.move_bcs()
        assert _same(_read(cl, f, f0.shape), ov.f), "move_bcs_PeriodicBC_VelocityInlet"
        d2q9.update_hydro_PeriodicBC_VelocityInlet(None, g2, lsz, f, u, v, rho, uw, ue, i32(nx), i32(ny)).wait()
        ov.update_hydro()
        for b, want in ((rho, ov.rho), (u, ov.u), (v, ov.v)):
            assert _same(_read(cl, b, (ny, nx)), want), "update_hydro_PeriodicBC_VelocityInlet"
        o.u[...] = ov.u
    # set_zero_velocity_in_obstacle (D2Q9.cl:377-396)
    d2q9.set_zero_velocity_in_obstacle(None, g2, lsz, m, u, v, i32(nx), i32(ny)).wait()
    zu = o.u.copy()
    zu[mask == 1] = 0
    assert _same(_read(cl, u, (ny, nx)), zu)


# ------------------------------------------------------------------------------------------------
# 2b. the periodic-box semantic (BASELINE configs 3 and 5) against the compiled
#     rocket_yeast.cl (float) and multicomponent_multiphase/multi.cl (double), one population
def _prog(cl, alias):
    try:
        return cl.Program.from_cache(cl.Context(), alias)
    except cl.Error as exc:
        pytest.skip(str(exc))


def test_periodic_streaming_equals_compiled_move_periodic(orc, cl):
    """`move_periodic` of rocket_yeast.cl:152-191 (float) and multi.cl:330-369 (double) == oracle_move_periodic,
    bit for bit: direction convention, wrap-around, nothing dropped."""
    from util import periodic_case
    nx, ny = 37, 21
    i32 = np.int32
    cx, cy = _buf(cl, CX), _buf(cl, CY)
    g2 = _gsize((nx, ny), (8, 4))
    for dtype, alias in ((np.float32, "rocket_yeast"), (np.float64, "multi")):
        prg = _prog(cl, alias)
        f0 = periodic_case(orc, nx, ny, dtype, amplitude=1e-2, seed=3)
        o = orc.OpenCLSchemeOracle(f0, 1.0, bc=orc.BC_PERIODIC, dtype=dtype)
        f, fs = _buf(cl, f0), _buf(cl, np.zeros_like(f0))
        for _ in range(3):
            if dtype == np.float32:
                prg.move_periodic(None, g2, (8, 4), f, fs, cx, cy, i32(nx), i32(ny), i32(1)).wait()
            else:
                prg.move_periodic(None, g2, (8, 4), f, fs, cx, cy, i32(nx), i32(ny), i32(0), i32(1), i32(9)).wait()
            f, fs = fs, f
            o.move()
            assert _same(_read(cl, f, f0.shape, dtype), o.f)


def test_periodic_fp64_step_agrees_with_compiled_multi_cl_to_rounding(orc, cl):
    """BASELINE config 5 / the fp64 gate: the oracle evaluates D2Q9.cl's formulas in double; the
    reference's only double-precision D2Q9 code is multi.cl (update_hydro_fluid :275-328, update_feq_fluid
    :11-75 D2Q9 branch, collide_particles_fluid :77-131 with zero body force).  Same algorithm, different
    association (loop-order sums, new_u/new_rho, pow(cs,4)): 300 steps of one against the other agree to
    1e-13 relative in rho and u -- two orders inside north_star's 1e-12."""
    from util import periodic_case
    prg = _prog(cl, "multi")
    nx, ny, omega = 48, 32, 1.7
    f0 = periodic_case(orc, nx, ny, np.float64, u0=0.05, amplitude=1e-3, seed=9)
    o = orc.OpenCLSchemeOracle(f0, omega, bc=orc.BC_PERIODIC, dtype=np.float64)
    i32, f64 = np.int32, np.float64
    Wd = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
    f, fs, feq = _buf(cl, f0), _buf(cl, f0), _buf(cl, np.zeros_like(f0))
    rho, u, v, gx, gy = (_buf(cl, np.zeros((ny, nx))) for _ in range(5))
    w, cx, cy = _buf(cl, Wd), _buf(cl, CX), _buf(cl, CY)
    cs = f64(1. / np.sqrt(3))
    g2, l2 = _gsize((nx, ny), (8, 8)), (8, 8)
    tail = (i32(nx), i32(ny), i32(0), i32(1), i32(9))
    steps = 300
    for _ in range(steps):           # multi.py:737-790: move, hydro, feq, collide
        prg.move_periodic(None, g2, l2, f, fs, cx, cy, *tail).wait()
        prg.copy_streamed_onto_f(None, g2, l2, fs, f, cx, cy, *tail).wait()
        prg.update_hydro_fluid(None, g2, l2, f, rho, u, v, gx, gy, w, cx, cy, *tail).wait()
        prg.update_feq_fluid(None, g2, l2, feq, rho, u, v, w, cx, cy, cs, *tail).wait()
        prg.collide_particles_fluid(None, g2, l2, f, feq, rho, u, v, gx, gy, f64(omega), w, cx, cy, *tail, cs).wait()
    o.run(steps)
    r_ref, u_ref, v_ref = (_read(cl, b, (ny, nx), np.float64) for b in (rho, u, v))
    assert np.isfinite(r_ref).all() and np.abs(u_ref).max() > 0.02
    assert np.abs(o.rho - r_ref).max() / np.abs(r_ref).max() < 1e-13
    assert np.abs(o.u - u_ref).max() / np.abs(u_ref).max() < 1e-13
    assert np.abs(o.v - v_ref).max() / np.abs(u_ref).max() < 1e-13
    assert np.abs(o.f - _read(cl, f, f0.shape, np.float64)).max() < 1e-14
    # mass: both conserve it to round-off
    assert abs(_read(cl, f, f0.shape, np.float64).sum() - f0.sum()) / f0.sum() < 1e-13


# ------------------------------------------------------------------------------------------------
# 3. the live host classes (build container only)
def test_live_opencl_reference_if_mounted(orc):
    from oracle import refload
    if not refload.opencl_host_available():
        pytest.skip("reference tree not mounted")
    m = refload.opencl_dim()
    np.random.seed(21)
    with refload.quiet() as out:
        sim = m.Pipe_Flow_Cylinder(cylinder_center=[0.6, 0.45], cylinder_radius=0.125, diameter=1., rho=1., viscosity=0.03,
                                   pressure_grad=-6., pipe_length=2.5, N=5, time_prefactor=0.15,
                                   two_d_local_size=(16, 4), three_d_local_size=(16, 4, 3))
    assert "omega" in out.getvalue()
    g0 = sim.get_fields()
    o = orc.OpenCLSchemeOracle(orc.from_opencl_host(g0["f"]), np.float32(sim.omega), np.float32(sim.inlet_rho),
                               np.float32(sim.outlet_rho), mask=orc.from_opencl_host(sim.obstacle_mask_host))
    assert 1.0 < sim.omega < 1.9 and sim.obstacle_mask_host.sum() > 20
    for n in (1, 30):
        sim.run(n)
        o.run(n)
        g = sim.get_fields()
        for k in ("f", "feq", "rho", "u", "v"):
            assert _same(orc.from_opencl_host(g[k]), getattr(o, k)), k
    assert np.isfinite(g["f"]).all() and np.abs(g["u"]).max() > 1e-3


def test_committed_golden_vectors_are_what_the_reference_produces_if_mounted():
    """Provenance: re-run the reference's own classes (same constructor calls, same seeds as
    tests/golden/make_golden.py) and compare with the committed files."""
    import ast
    from oracle import refload
    if not refload.opencl_host_available():
        pytest.skip("reference tree not mounted")
    ls = dict(two_d_local_size=(16, 16), three_d_local_size=(16, 16, 1))    # another work-group size on purpose
    for name, mod, cls in (("opencl_pipe_65x33.npz", refload.opencl_dim(), "Pipe_Flow"),
                           ("opencl_d2q9i_pipe_49x25.npz", refload.opencl_dim_D2Q9i(), "Pipe_Flow"),
                           ("oldcl_velocity_inlet_61x31.npz", refload.old_opencl(), "Pipe_Flow_PeriodicBC_VelocityInlet")):
        g = _load(name)
        kw = ast.literal_eval(str(g["ctor_kwargs"]))
        np.random.seed(int(g["seed"]))
        with refload.quiet():
            sim = getattr(mod, cls)(**kw, **ls)
        get = sim.get_fields if hasattr(sim, "get_fields") else sim.get_fields_on_cpu
        assert _same(get()["f"], g["f_0"]), name
        s = int(g["steps"][1])
        sim.run(s)
        got = get()
        for k in ("f", "rho", "u", "v"):
            assert _same(got[k], g[f"{k}_{s}"]), (name, k)


def test_live_single_stage_methods_if_mounted(orc):
    """The check notebooks call move(), move_bcs(), ... one by one (testing/Bryan/opencl_check_03.ipynb)."""
    from oracle import refload
    if not refload.opencl_host_available():
        pytest.skip("reference tree not mounted")
    m = refload.opencl_dim()
    np.random.seed(22)
    with refload.quiet():
        sim = m.Pipe_Flow(diameter=1., rho=1., viscosity=0.1, pressure_grad=-1., pipe_length=1.5, N=20, time_prefactor=3.,
                          two_d_local_size=(8, 8), three_d_local_size=(8, 8, 1))
    o = orc.OpenCLSchemeOracle(orc.from_opencl_host(sim.get_fields()["f"]), np.float32(sim.omega),
                               np.float32(sim.inlet_rho), np.float32(sim.outlet_rho))
    for stage in ("move", "move_bcs", "update_hydro", "update_feq", "collide_particles", "move", "move_bcs"):
        getattr(sim, stage)()
        getattr(o, stage)()
        g = sim.get_fields()
        assert _same(orc.from_opencl_host(g["f"]), o.f), stage
    assert _same(orc.from_opencl_host(g["u"]), o.u)


def test_the_references_own_cross_check_if_mounted():
    """testing/Bryan/opencl_check_03.ipynb builds a Cython sim and an OpenCL sim with identical arguments,
    calls one method on each and looks at |a-b| > 1e-6; its notes say "the corners are definitely different
    though.  And the walls."  Both implementations run here, so the procedure can be repeated: identical
    initial populations, interiors that agree to 1e-6 after the first boundary pass and the first
    streaming pass, and differences confined to boundary nodes -- the reference's two implementations are
    two algorithms, which is why this repo carries one scheme (and one oracle) per implementation."""
    from oracle import refload
    if not (refload.opencl_host_available() and refload.available()):
        pytest.skip("reference tree not mounted")
    oc, ol = refload.old_cython(), refload.old_opencl()
    kw = dict(omega=1.0, lx=48, ly=24, deltaP=-0.01)
    np.random.seed(0)
    a = oc.Pipe_Flow(**kw)
    np.random.seed(0)
    with refload.quiet():
        b = ol.Pipe_Flow(**kw, two_d_local_size=(8, 8), three_d_local_size=(8, 8, 1))

    def both():
        return np.asarray(a.f), np.rollaxis(b.get_fields_on_cpu()["f"], 2, 0)      # cell 34's `check_variable`

    fa, fb = both()
    assert np.array_equal(fa, fb)
    for stage, rim in (("move_bcs", 1), ("move", 2)):        # streaming carries a boundary difference one node inwards
        getattr(a, stage)()
        getattr(b, stage)()
        fa, fb = both()
        differ = (np.abs(fa - fb) > 1e-6).any(axis=0)                                 # (nx, ny)
        assert differ.any(), stage                                                    # they do disagree ...
        assert not differ[rim:-rim, rim:-rim].any(), stage                            # ... next to the boundary only


# ------------------------------------------------------------------------------------------------
# 4. the emulation layer
TOY = """
__kernel void ids(__global int *out, const int nx, const int ny)
{
    const int x = get_global_id(0), y = get_global_id(1);
    if (x < nx && y < ny)
        out[y*nx + x] = 1000000*get_group_id(1) + 10000*get_group_id(0) + 100*get_local_id(1) + get_local_id(0);
}
// reverses each work-group's slice through local memory: wrong without a working barrier
__kernel void reverse_in_group(__global const float *src, __global float *dst, __local float *tile)
{
    const int l = get_local_id(0), n = get_local_size(0), g = get_global_id(0);
    tile[l] = src[g];
    barrier(CLK_LOCAL_MEM_FENCE);
    float mine = tile[n - 1 - l];
    barrier(CLK_LOCAL_MEM_FENCE);
    tile[l] = 2.f*mine;                      /* second round: barrier count > 1 */
    barrier(CLK_LOCAL_MEM_FENCE);
    dst[g] = tile[l] + get_num_groups(0);
}
__kernel void promote(__global float *out, const float a, const double b)
{
    out[0] = a*(2./3.);                      /* double literal: evaluated in double, rounded on store */
    out[1] = a*(2.f/3.f);
    out[2] = (float)(b*a);
}
"""


def test_clshim_ndrange_ids_and_barriers(cl):
    prg = cl.Program(cl.Context(), TOY).build()
    nx, ny = 10, 6
    out = _buf(cl, np.full((ny, nx), -1, np.int32))
    prg.ids(None, (12, 6), (4, 3), out, np.int32(nx), np.int32(ny)).wait()
    got = _read(cl, out, (ny, nx), np.int32)
    y, x = np.mgrid[0:ny, 0:nx]
    assert np.array_equal(got, 1000000 * (y // 3) + 10000 * (x // 4) + 100 * (y % 3) + (x % 4))

    src = np.arange(64, dtype=np.float32)
    dst = _buf(cl, np.zeros(64, np.float32))
    prg.reverse_in_group(None, (64,), (16,), _buf(cl, src), dst, cl.LocalMemory(4 * 16)).wait()
    want = 2 * src.reshape(4, 16)[:, ::-1].reshape(-1) + 4
    assert np.array_equal(_read(cl, dst, (64,)), want)


def test_clshim_arithmetic_promotion_and_argument_checks(cl):
    prg = cl.Program(cl.Context(), TOY).build()
    out = _buf(cl, np.zeros(3, np.float32))
    a = np.float32(0.1234567)
    prg.promote(None, (1,), (1,), out, a, np.float64(3.0)).wait()
    got = _read(cl, out, (3,))
    assert got[0] == np.float32(np.float64(a) * (2. / 3.))
    assert got[1] == a * (np.float32(2) / np.float32(3))
    assert got[2] == np.float32(3.0 * np.float64(a))
    with pytest.raises(cl.LogicError):
        prg.promote(None, (1,), (1,), out, 0.5, np.float64(3.0))            # unsized Python float
    with pytest.raises(cl.LogicError):
        prg.promote(None, (1,), (1,), out, np.float64(0.5), np.float64(3.0))  # 8 bytes for a float
    with pytest.raises(cl.LogicError):
        prg.promote(None, (1,), (1,), out, a)                                  # missing argument
    with pytest.raises(cl.RuntimeError):
        prg.ids(None, (10, 6), (4, 3), out, np.int32(1), np.int32(1))          # 10 % 4 != 0
    with pytest.raises(cl.RuntimeError):
        cl.Program(cl.Context(), "__kernel void broken(__global float *x) { x[0] = ; }").build()


def test_clshim_buffers_keep_the_host_memory_order(cl):
    a = np.asfortranarray(np.arange(12, dtype=np.float32).reshape(3, 4))
    b = _buf(cl, a) if False else cl.Buffer(None, cl.mem_flags.COPY_HOST_PTR, hostbuf=a)
    flat = np.empty(12, np.float32)
    cl.enqueue_copy(None, flat, b)
    assert np.array_equal(flat, a.reshape(-1, order="F"))
    back = np.zeros((3, 4), np.float32, order="F")
    cl.enqueue_copy(None, back, b)
    assert np.array_equal(back, a)
