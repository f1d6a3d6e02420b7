"""The oracle itself is checked before it is trusted (runs on CPU, `-m "not gpu"`).

1. scheme 'cython' restatement  == committed golden vectors, BIT FOR BIT.  The vectors were
   produced by the unmodified reference (tests/golden/make_golden.py -> compiled
   LB_D2Q9/dimensionless/cython_dim.pyx and LB_D2Q9/OLD/cython.pyx).
2. scheme 'opencl' restatement (the path being replaced) is pinned bit for bit against the
   reference's own kernels in tests/test_opencl_reference.py; here: (a) the interior update it
   shares with the Cython path, on the golden vectors, (b) the reference's Poiseuille known answer
   and stored constructor printouts (docs/opencl_dimensionless_verification.ipynb), (c) internal
   consistency properties.
"""
import os

import numpy as np
import pytest

from util import periodic_case, pipe_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _yx(a):
    """(9,nx,ny)/(nx,ny) reference host arrays -> ([9,]ny,nx) device layout."""
    return np.ascontiguousarray(a.transpose(0, 2, 1) if a.ndim == 3 else a.T)


@pytest.mark.parametrize("name", ["cython_pipe_65x33.npz", "cython_cylinder_121x41.npz", "old_obstacles_49x25.npz",
                                  "old_velocity_inlet_61x31.npz", "old_velocity_inlet_obstacles_61x31.npz"])
def test_cython_scheme_matches_reference_golden_bitexact(orc, name):
    g = _load(name)
    mask = _yx(g["mask"]) if "mask" in g.files else None
    vin = (float(g["u_w"]), float(g["u_e"])) if "u_w" in g.files else None
    o = orc.CythonSchemeOracle(_yx(g["f_0"]), _yx(g["u_0"]), _yx(g["v_0"]), float(g["omega"]),
                               float(g["inlet_rho"]), float(g["outlet_rho"]), mask=mask,
                               old_api=name.startswith("old_"), velocity_inlet=vin)
    done = 0
    for s in g["steps"]:
        o.run(int(s) - done)
        done = int(s)
        assert np.array_equal(o.f, _yx(g[f"f_{s}"])), f"f after {s} steps"
        assert np.array_equal(o.rho, _yx(g[f"rho_{s}"])), f"rho after {s} steps"
        assert np.array_equal(o.u, _yx(g[f"u_{s}"])), f"u after {s} steps"
        assert np.array_equal(o.v, _yx(g[f"v_{s}"])), f"v after {s} steps"


def test_live_reference_if_present(orc):
    """Where oracle/_ref exists (build container and GPU box), re-run the reference itself."""
    from oracle import refload
    if not refload.available():
        pytest.skip("oracle/_ref not built")
    cd = refload.cython_dim()
    np.random.seed(7)
    with refload.quiet():
        sim = cd.Pipe_Flow(diameter=1., rho=1., viscosity=0.05, pressure_grad=-1., pipe_length=1.5, N=24, time_prefactor=4.)
    o = orc.CythonSchemeOracle(orc.from_cython_host(sim.f), orc.from_cython_host(sim.u), orc.from_cython_host(sim.v),
                               sim.omega, sim.inlet_rho, sim.outlet_rho)
    sim.run(60)
    o.run(60)
    assert np.array_equal(orc.from_cython_host(sim.f), o.f)
    assert np.array_equal(orc.from_cython_host(sim.u), o.u)


@pytest.mark.parametrize("name", ["cython_pipe_65x33.npz", "cython_cylinder_121x41.npz"])
def test_opencl_scheme_interior_agrees_with_reference(orc, name):
    """One step from the reference's own initial populations: away from the boundary (and from the
    obstacle's zeroed velocity) the OpenCL-order and Cython-order schemes are the same algorithm --
    stream, moments, equilibrium, BGK -- so the float32 oracle must reproduce the reference's f to
    float32 rounding there."""
    g = _load(name)
    f0 = _yx(g["f_0"])
    mask = _yx(g["mask"]).astype(bool) if "mask" in g.files else np.zeros(f0.shape[1:], bool)
    o = orc.OpenCLSchemeOracle(f0, float(g["omega"]), float(g["inlet_rho"]), float(g["outlet_rho"]),
                               mask=mask.astype(np.int32) if mask.any() else None)
    o.run(1)
    want = _yx(g["f_1"])
    inner = np.zeros(mask.shape, bool)
    inner[2:-2, 2:-2] = True
    # exclude solid nodes and their neighbours: the Cython class zeroes u,v inside the obstacle
    # every step (cython_dim.pyx:459-466), opencl_dim does not
    near = mask.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            near |= np.roll(np.roll(mask, dy, 0), dx, 1)
    sel = inner & ~near
    assert sel.sum() > 0.5 * sel.size
    err = np.abs(o.f[:, sel] - want[:, sel]).max()
    assert err <= 5e-7, err          # a few float32 ulps of f (up to ~0.85 here); the init noise is 1e-3*f


def test_opencl_scheme_zero_velocity_option_matches_cython_obstacle_interior(orc):
    """With zero_obstacle_velocity the solid nodes behave like the Cython class's: compare one step
    on nodes inside the obstacle (bounce-back + collision with u=v=0)."""
    g = _load("cython_cylinder_121x41.npz")
    f0, mask = _yx(g["f_0"]), _yx(g["mask"]).astype(bool)
    o = orc.OpenCLSchemeOracle(f0, float(g["omega"]), float(g["inlet_rho"]), float(g["outlet_rho"]),
                               mask=mask.astype(np.int32), zero_obstacle_velocity=True)
    o.run(1)
    assert (o.u[mask] == 0).all() and (o.v[mask] == 0).all()
    # rho on solid nodes is a plain sum of the (permuted) streamed populations in both schemes ...
    core = mask & np.roll(mask, 1, 0) & np.roll(mask, -1, 0) & np.roll(mask, 1, 1) & np.roll(mask, -1, 1)
    assert core.any()
    assert np.abs(o.rho[core] - _yx(g["rho_1"])[core]).max() <= 5e-7


def test_reference_constructor_printouts(orc):
    """Golden parameter values stored in the reference's notebooks:
    opencl_dimensionless_verification.ipynb cells 7,9,10: omega 0.324465802203, inlet rho 1.063 /
    1.002424 / 1.000150375 for N=10/50/200; python_cython_opencl_comparison.ipynb cells 10-12:
    omega 0.413223140496, inlet rho 1.00368738304, grid 3751x1251 for the N=125 cylinder."""
    import lb_b200.dimensionless as lb

    class HostOnly(lb.Pipe_Flow):             # host algebra only: no device objects
        def init_cuda(self):
            pass

        def init_hydro(self):
            self._set_boundary_densities()

        def update_feq(self):
            pass

        def init_pop(self):
            pass

    class HostOnlyCyl(lb.Pipe_Flow_Cylinder, HostOnly):
        init_hydro = HostOnly.init_hydro

    for N, want_rho in ((10, 1.063), (50, 1.002424), (200, 1.000150375)):
        s = HostOnly(diameter=1.5, rho=10., viscosity=5., pressure_grad=-100., pipe_length=3., N=N, verbose=False)
        assert abs(s.omega - 0.324465802203) < 5e-13
        assert abs(s.inlet_rho - want_rho) < 5e-10
        assert (s.nx, s.ny) == (2 * N + 1, N + 1)
    kw = dict(cylinder_center=[3. / 4, .5], cylinder_radius=.1, diameter=1., rho=1., viscosity=1., pressure_grad=-10.,
              pipe_length=3., N=125, verbose=False)
    c = HostOnlyCyl(**kw)
    assert (c.nx, c.ny) == (3751, 1251)
    assert c.two_d_global_size == (3776, 1280) and c.three_d_global_size == (3776, 1280, 9)
    assert c.obstacle_mask_host.sum() > 0 and c.obstacle_mask_host.dtype == np.int32
    # the benchmark notebook's printouts ("Reynolds number: 1.5625", omega 0.413223140496, inlet rho
    # 1.00368738304) were produced with the Cython-style algebra (an older opencl_dim revision):
    c = HostOnlyCyl(units="cython", **kw)
    assert abs(c.L - 0.1) < 1e-15 and abs(c.T - 0.08) < 1e-15 and abs(c.Re - 1.5625) < 1e-12
    assert abs(c.omega - 0.413223140496) < 5e-13
    assert abs(c.inlet_rho - 1.00368738304) < 5e-12
    # docs/cs205_movie.ipynb cell 7: omega 1.92604006163, inlet rho 1.009228288, Re 156.25, 751x251
    m = HostOnlyCyl(units="cython", cylinder_center=[3. / 4, .5], cylinder_radius=.1, diameter=1., rho=1., viscosity=1.,
                    pressure_grad=-100., pipe_length=3., N=25, verbose=False)
    assert abs(m.Re - 156.25) < 1e-9 and abs(m.omega - 1.92604006163) < 5e-12
    assert abs(m.inlet_rho - 1.009228288) < 5e-10 and (m.nx, m.ny) == (751, 251)


def test_poiseuille_known_answer_on_oracle(orc):
    """docs/opencl_dimensionless_verification.ipynb (N=10, 999 steps): x-averaged physical u(y)
    against u = (1/(2 rho nu)) grad_p y (y - D).  Peak 0.5625 analytically; the N=10 lattice
    overshoots to ~0.572 (visible in pictures/resolution_convergence.png); RMS error ~7e-3."""
    D, rho_p, nu, gp, N = 1.5, 10., 5., -100., 10
    L, T = D, np.sqrt(D / (abs(gp) / rho_p))
    W = (abs(gp) / rho_p) * L * T / nu
    dx = 1. / N
    dt = dx ** 2
    omega = 1. / (3 * (dt / dx ** 2) / W + 0.5)
    nx, ny = 2 * N + 1, N + 1
    rin = 1. + nx * (dt ** 2 / dx) * (1. / orc.cs2)
    f0, _ = pipe_case(orc, nx, ny, np.float32, inlet_rho=rin, seed=0)
    for dtype in (np.float32, np.float64):
        o = orc.OpenCLSchemeOracle(f0, omega, rin, 1.0, dtype=dtype)
        o.run(int(10. / dt))
        u_phys = o.u.astype(np.float64) * (dx / dt) * (L / T)
        prof = u_phys.mean(axis=1)
        y = np.linspace(0, D, ny)
        theory = (1. / (2 * rho_p * nu)) * gp * y * (y - D)
        assert abs(prof.max() - 0.5625) < 0.015
        assert np.sqrt(np.mean((prof - theory) ** 2)) < 0.01
        assert abs(prof[0]) < 0.02 and abs(prof[-1]) < 0.02


def test_stage_sequence_equals_run_and_is_deterministic(orc):
    f0, m = pipe_case(orc, 64, 40, np.float32, mask="blocks")
    a = orc.OpenCLSchemeOracle(f0, 1.2, 1.01, 1.0, mask=m)
    b = orc.OpenCLSchemeOracle(f0, 1.2, 1.01, 1.0, mask=m)
    a.run(6)
    for _ in range(6):
        b.move(); b.move_bcs(); b.update_hydro(); b.update_feq(); b.collide_particles()
    assert np.array_equal(a.f, b.f) and np.array_equal(a.u, b.u)


def test_stale_stream_slots_never_reach_the_result(orc):
    """the push kernel leaves slots with no upstream node untouched; every one of them
    is overwritten by move_bcs.  Poison the second buffer with NaN and check nothing leaks."""
    f0, m = pipe_case(orc, 48, 30, np.float32, mask="touching")
    o = orc.OpenCLSchemeOracle(f0, 1.2, 1.01, 1.0, mask=m)
    o.f_streamed[:] = np.nan
    o.run(20)
    assert np.isfinite(o.f).all() and np.isfinite(o.rho).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_periodic_box_conserves_mass_and_momentum(orc, dtype):
    f0 = periodic_case(orc, 96, 64, dtype, amplitude=1e-3)
    o = orc.OpenCLSchemeOracle(f0, 1.7, bc=orc.BC_PERIODIC, dtype=dtype)
    m0 = f0.astype(np.float64).sum()
    px0 = (f0.astype(np.float64) * orc.CX[:, None, None]).sum()
    o.run(200)
    tol = 1e-5 if dtype == np.float32 else 1e-13     # fp32 round-off random-walks at ~1e-8 per step
    assert abs(o.f.astype(np.float64).sum() - m0) / m0 < tol
    assert abs((o.f.astype(np.float64) * orc.CX[:, None, None]).sum() - px0) / m0 < tol


def test_periodic_translation_invariance(orc):
    """Shifting the initial condition of a periodic box shifts the result, bit for bit."""
    f0 = periodic_case(orc, 40, 24, np.float64, amplitude=1e-3)
    a = orc.OpenCLSchemeOracle(f0, 1.5, bc=orc.BC_PERIODIC, dtype=np.float64)
    b = orc.OpenCLSchemeOracle(np.roll(f0, (5, 11), axis=(1, 2)), 1.5, bc=orc.BC_PERIODIC, dtype=np.float64)
    a.run(30)
    b.run(30)
    assert np.array_equal(np.roll(a.f, (5, 11), axis=(1, 2)), b.f)


def test_f32_and_f64_oracles_agree_to_single_precision(orc):
    f0, m = pipe_case(orc, 64, 40, np.float32, mask="blocks")
    a = orc.OpenCLSchemeOracle(f0, 1.2, 1.01, 1.0, mask=m, dtype=np.float32)
    b = orc.OpenCLSchemeOracle(f0, 1.2, 1.01, 1.0, mask=m, dtype=np.float64)
    a.run(50)
    b.run(50)
    assert np.abs(a.rho - b.rho).max() < 5e-6


def test_d2q9i_equilibrium_as_shipped_sums_to_rho_squared(orc):
    """Documents the reference's incompressible variant as it is (D2Q9i.cl:58-59):
    feq_j = w_j * rho * (rho + 3 c.u + 4.5 (c.u)^2 - 1.5 u^2), whose zeroth moment is rho^2, not
    rho -- so rho = 1 is an unstable fixed point of its collision and long runs diverge.  The
    restatement is literal and pinned to the compiled D2Q9i.cl (tests/test_opencl_reference.py)."""
    ny, nx = 8, 12
    rho = np.full((ny, nx), 1.05)
    u = np.full((ny, nx), 0.02)
    v = np.full((ny, nx), -0.01)
    feq = orc.feq_of(rho, u, v, np.float64, incompressible=True)
    usq = 0.02 ** 2 + 0.01 ** 2
    assert np.allclose(feq.sum(axis=0), 1.05 * (1.05 + 3 * usq - 1.5 * usq) , rtol=0, atol=1e-12) or \
        np.allclose(feq.sum(axis=0), 1.05 * 1.05, atol=2e-3)
    mom_x = (feq * orc.CX[:, None, None]).sum(axis=0)
    assert np.allclose(mom_x, 1.05 * 0.02, atol=1e-12)           # first moment = rho * u (u is the raw momentum)
    std = orc.feq_of(rho, u, v, np.float64)
    assert np.allclose(std.sum(axis=0), 1.05, atol=1e-12)        # the compressible equilibrium conserves mass
