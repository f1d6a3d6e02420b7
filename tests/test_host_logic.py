"""Host-side logic that needs no GPU: slab partitioning, the multi-process rendezvous under gloo
(world_size 2), mask rasterisation, layout conventions."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_slabs_and_edges():
    from lb_b200.lattice import slab_edges, split_slabs
    assert split_slabs(32768, 8) == [(i * 4096, 4096) for i in range(8)]
    r = split_slabs(203, 5)
    assert sum(w for _, w in r) == 203 and r[0] == (0, 41) and r[-1] == (163, 40)
    assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(4))
    with pytest.raises(ValueError):
        split_slabs(7, 4)
    assert slab_edges(0, 1, "pipe") == ("boundary", "boundary")
    assert slab_edges(0, 1, "periodic") == ("wrap", "wrap")
    assert slab_edges(0, 4, "pipe") == ("boundary", "halo")
    assert slab_edges(2, 4, "pipe") == ("halo", "halo")
    assert slab_edges(3, 4, "pipe") == ("halo", "boundary")
    assert slab_edges(3, 4, "periodic") == ("halo", "halo")


def test_fortran_host_arrays_are_device_layout():
    """opencl_dim's (nx,ny,9) Fortran-order arrays are byte-identical to f[9][ny][nx]."""
    nx, ny = 5, 3
    a = np.zeros((nx, ny, 9), dtype=np.float32, order="F")
    for j in range(9):
        for y in range(ny):
            for x in range(nx):
                a[x, y, j] = j * 100 + y * 10 + x
    dev = a.T
    assert dev.flags.c_contiguous and dev.shape == (9, ny, nx)
    assert np.shares_memory(dev, a)
    assert dev[4, 2, 3] == 423
    flat = np.frombuffer(a.tobytes(order="A"), dtype=np.float32)
    assert flat[4 * nx * ny + 2 * nx + 3] == 423          # jump_id*nx*ny + y*nx + x  (D2Q9.cl:24-25)


def test_circle_matches_skimage_contract():
    from lb_b200 import draw
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
    from skimage.draw import circle as shim_circle          # independent statement of the contract
    for (r, c, R) in ((30.0, 20.0, 4), (193.75, 162.5, 125), (40.5, 33.25, 10)):
        a = draw.circle(r, c, R)
        b = shim_circle(r, c, R)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        m = np.zeros((int(r + R + 2), int(c + R + 2)), bool)
        m[a[0], a[1]] = True
        assert abs(m.sum() - np.pi * R * R) / (np.pi * R * R) < 0.12


class _FakeLattice:
    """Records what SlabLattice asks of the per-rank lattice."""

    def __init__(self, nx, ny, omega, inlet_rho, outlet_rho, **kw):
        self.nx, self.ny, self.kw = nx, ny, kw
        self.connected = {}
        self.primed = 0
        self.f = None

    def halo_ipc_handle(self):
        return (b"H%03d" % self.kw["x_offset"]).ljust(64, b"\0")

    def halo_connect_ipc(self, side, handle, device):
        self.connected[side] = (handle.rstrip(b"\0"), device)

    def halo_prime(self):
        self.primed += 1

    def sync(self):
        pass

    def upload_f(self, f):
        self.f = f

    def download(self, field):
        return self.f

    def close(self):
        pass


def _worker(rank, world, port, bc, out):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "2d-lb_b200"))
    from lb_b200.slab import SlabLattice
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        s = SlabLattice(37, 6, 1.0, bc=bc, lattice_factory=_FakeLattice, device=rank)
        f = np.arange(9 * 6 * 37, dtype=np.float32).reshape(9, 6, 37)
        s.upload_f(f)
        whole = s.gather("f")
        out.put((rank, s.x_offset, s.nx, s.west_edge, s.east_edge, dict(s.lat.connected), s.lat.primed,
                 bool(np.array_equal(whole, f))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bc", ["pipe", "periodic"])
def test_slab_rendezvous_gloo_world2(bc):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bc, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert (r0[1], r0[2]) == (0, 19) and (r1[1], r1[2]) == (19, 18)
    assert r0[7] and r1[7]                      # slab slices reassemble the global array
    assert r0[6] == 1 and r1[6] == 1            # primed once after upload
    if bc == "pipe":
        assert (r0[3], r0[4]) == ("boundary", "halo") and (r1[3], r1[4]) == ("halo", "boundary")
        assert r0[5] == {"east": (b"H019", 1)} and r1[5] == {"west": (b"H000", 0)}
    else:
        assert r0[5] == {"west": (b"H019", 1), "east": (b"H019", 1)}
        assert r1[5] == {"west": (b"H000", 0), "east": (b"H000", 0)}


def test_mask_ingestion(tmp_path):
    """lb_b200.masks: image -> (nx, ny) bool, nearest-neighbour resampling, bit packing."""
    from PIL import Image
    from lb_b200 import masks
    img = np.zeros((40, 80), np.uint8)            # (H, W), like a TIFF read by PIL
    img[10:20, 30:50] = 255
    path = tmp_path / "m.png"
    Image.fromarray(img).save(path)
    m = masks.from_image(str(path))
    assert m.shape == (80, 40) and m.dtype == bool
    assert m[30:50, 10:20].all() and m.sum() == 200
    r = masks.resample(m, 160, 80)
    assert r.shape == (160, 80) and r.sum() == 800 and r[60:100, 20:40].all()
    assert np.array_equal(masks.resample(m, 80, 40), m)
    assert np.array_equal(masks.unpack(masks.pack(m)), m)
    # anti-aliased route (grey resize, then threshold): identity at the same size, area-exact when shrinking by an
    # integer factor, mean-preserving, and equal to nearest neighbour on a mask whose edges fall on the coarse grid
    grey = m.astype(np.float64)
    assert np.array_equal(masks.resize(grey, 80, 40), grey)
    half = masks.resize(grey, 40, 20)
    assert half.shape == (40, 20) and np.allclose(half, grey.reshape(40, 2, 20, 2).mean(axis=(1, 3)))
    third = masks.resize(grey, 27, 13)
    assert abs(third.mean() - grey.mean()) < 1e-12 and third.min() >= 0 and third.max() <= 1 + 1e-12
    up = masks.resize(grey, 160, 80)
    assert up.shape == (160, 80) and abs(up.mean() - grey.mean()) < 2e-3 and up[70:90, 25:35].min() == 1.0
    aa = masks.from_image(str(path), 40, 20, antialias=True)
    assert np.array_equal(aa, masks.from_image(str(path), 40, 20))
    img2 = np.zeros((40, 80), np.uint8)
    img2[11:20, 31:50] = 255                          # odd edges: the coarse cells on the rim are half covered
    Image.fromarray(img2).save(tmp_path / "m2.png")
    a2 = masks.from_image(str(tmp_path / "m2.png"), 40, 20, antialias=True, threshold=0.4)
    # 36 fully covered coarse cells + 4 + 9 half covered ones on the two rims (>= 40 %: solid); the corner (25 %) is not
    assert a2.sum() == 49 and a2[16:25, 6:10].all() and a2[15, 6:10].all() and a2[16:25, 5].all() and not a2[15, 5]
    g = np.load(os.path.join(ROOT, "tests", "golden", "cs205_binary_mask.npz"))
    src = masks.unpack(g)
    assert src.shape == (800, 400) and abs(src.mean() - 0.087) < 0.001      # 8.7 % solid
    assert not (src[0].any() or src[-1].any() or src[:, 0].any() or src[:, -1].any())


@pytest.mark.parametrize("name", ["opencl_pipe_65x33", "opencl_cylinder_121x41", "opencl_d2q9i_pipe_49x25",
                                  "opencl_d2q9i_cylinder_121x41"])
def test_class_parameter_algebra_equals_the_reference_opencl_classes(name):
    """Grid size, omega, inlet density and cylinder mask computed by the drop-in classes (device calls
    stubbed out) equal what the reference's opencl_dim / opencl_dim_D2Q9i classes computed for the
    same constructor call (stored with the golden vectors), to the last bit."""
    import ast
    import lb_b200.dimensionless as lb
    from lb_b200 import dimensionless_D2Q9i as lbi
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    kw = ast.literal_eval(str(g["ctor_kwargs"]))
    mod = lbi if "d2q9i" in name else lb
    base = mod.Pipe_Flow_Cylinder if "cylinder" in name else mod.Pipe_Flow

    class NoDevice(base):
        def init_cuda(self):
            pass

        def init_hydro(self):
            self._set_boundary_densities()

        def update_feq(self):
            pass

        def init_pop(self):
            pass

    s = NoDevice(verbose=False, **kw)
    assert (s.nx, s.ny) == (int(g["nx"]), int(g["ny"]))
    assert float(s.omega) == float(g["omega"]) and float(s.inlet_rho) == float(g["inlet_rho"])
    assert float(s.outlet_rho) == float(g["outlet_rho"])
    if "mask" in g.files:
        assert np.array_equal(s.obstacle_mask_host, g["mask"])


def _tb2_host():
    """tools/libtb2_host.so: the two phases of the temporally blocked kernel compiled for the host."""
    import ctypes as ct
    import shutil
    import subprocess
    so = os.path.join(ROOT, "tools", "libtb2_host.so")
    src = os.path.join(ROOT, "tools", "tb2_host.cu")
    deps = [src] + [os.path.join(ROOT, "2d-lb_b200", "csrc", n) for n in ("lb_tb2.cuh", "lb_device.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        if not shutil.which("nvcc"):
            pytest.skip("nvcc not available to build the host replay")
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--shared",
                        "-Xcompiler", "-fPIC", "-o", so, src], check=True, capture_output=True)
    return ct.CDLL(so)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_temporal_blocking_tile_logic_on_cpu(dtype):
    """csrc/lb_tb2.cuh's phases are host-callable: every CTA is replayed on the CPU (shared memory
    poisoned with NaN first) and the result must equal TWO steps of the oracle bit for bit -- pipe flow
    with obstacles on every edge, periodic boxes down to 2 x 2, five tile shapes, odd thread counts."""
    import ctypes as ct
    from oracle import oracle as orc
    from util import periodic_case, pipe_case
    orc.build()
    L = _tb2_host()

    def replay(f0, mask, bc, shape, nthreads, omega, zero_vel=0):
        _, ny, nx = f0.shape
        out = np.full_like(f0, np.nan)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = L.tb2_host_run(nx, ny, bc, int(dtype == np.float64), f0.ctypes.data_as(ct.c_void_p),
                            out.ctypes.data_as(ct.c_void_p), None if m is None else m.ctypes.data_as(ct.c_void_p),
                            ct.c_double(omega), ct.c_double(1.01), ct.c_double(1.0), ct.c_double(orc.cs2),
                            ct.c_double(orc.cs22), ct.c_double(orc.two_cs4), zero_vel, shape, nthreads)
        assert rc == 0
        return out

    for (nx, ny) in ((97, 41), (256, 34), (5, 4), (2, 2), (129, 17)):
        for maskk in ("none", "touching"):
            f0, m = pipe_case(orc, nx, ny, dtype, mask=maskk if nx > 8 else "none", seed=nx)
            for zv in ((0, 1) if m is not None else (0,)):
                ref = orc.OpenCLSchemeOracle(f0, 1.3, 1.01, 1.0, mask=m, dtype=dtype, zero_obstacle_velocity=bool(zv))
                ref.run(2)
                for shape, nt in ((0, 256), (1, 64), (2, 33), (3, 7), (4, 1)):
                    assert np.array_equal(replay(f0, m, 0, shape, nt, 1.3, zv), ref.f), (nx, ny, maskk, zv, shape)
    for (nx, ny) in ((96, 40), (130, 67), (7, 5), (3, 3), (2, 2)):
        f0 = periodic_case(orc, nx, ny, dtype, amplitude=1e-3, seed=ny)
        ref = orc.OpenCLSchemeOracle(f0, 1.7, bc=orc.BC_PERIODIC, dtype=dtype)
        ref.run(2)
        for shape, nt in ((0, 256), (1, 64), (3, 5), (4, 2)):
            assert np.array_equal(replay(f0, None, 1, shape, nt, 1.7), ref.f), (nx, ny, shape)


@pytest.mark.parametrize("parts", [2, 3])
def test_two_update_halo_protocol_is_sufficient(parts):
    """Design check for the next step of the multi-GPU path: a slab that advances TWO
    updates per exchange needs, per interior face and row, populations {0,2,4} and the three incoming ones of
    the neighbour's boundary column plus the three incoming ones of the column behind it (and the mask of the
    boundary column) -- nothing else.  Model: every slab is advanced two oracle steps on its own columns
    extended by two ghost columns that hold exactly those values and NaN everywhere else; its own columns
    must come out NaN-free and bit-identical to the undivided lattice."""
    from oracle import oracle as orc
    from util import pipe_case
    orc.build()
    nx, ny = 96, 33
    f0, mask = pipe_case(orc, nx, ny, np.float32, mask="touching", seed=5)
    mask[ny // 2 - 3: ny // 2 + 3, nx // parts - 2: nx // parts + 3] = 1          # a body across the first cut
    whole = orc.OpenCLSchemeOracle(f0, 1.3, 1.01, 1.0, mask=mask)
    cuts = [round(k * nx / parts) for k in range(parts + 1)]
    slabs = [f0[:, :, cuts[k]:cuts[k + 1]].copy() for k in range(parts)]
    east_in, west_in = (1, 5, 8), (3, 6, 7)                                          # populations moving +x / -x
    for _ in range(3):                                                               # three exchanges = six updates
        whole.run(2)
        new = []
        for k, own in enumerate(slabs):
            x0, x1 = cuts[k], cuts[k + 1]
            gw = 2 if k > 0 else 0
            ge = 2 if k < parts - 1 else 0
            ext = np.full((9, ny, gw + (x1 - x0) + ge), np.nan, np.float32)
            ext[:, :, gw:gw + (x1 - x0)] = own
            if gw:                                                                   # from the west neighbour
                nb = slabs[k - 1]
                for j in (0, 2, 4) + east_in:
                    ext[j, :, 1] = nb[j, :, -1]                                      # its boundary column
                for j in east_in:
                    ext[j, :, 0] = nb[j, :, -2]                                      # the column behind it
            if ge:
                nb = slabs[k + 1]
                for j in (0, 2, 4) + west_in:
                    ext[j, :, -2] = nb[j, :, 0]
                for j in west_in:
                    ext[j, :, -1] = nb[j, :, 1]
            m = np.zeros((ny, ext.shape[2]), np.uint8)
            lo, hi = x0 - min(gw, 1), x1 + min(ge, 1)                                 # mask: own columns + one ghost column
            m[:, gw - min(gw, 1): gw + (x1 - x0) + min(ge, 1)] = mask[:, lo:hi]
            sub = orc.OpenCLSchemeOracle(ext, 1.3, 1.01, 1.0, mask=m)
            with np.errstate(all="ignore"):
                sub.run(2)
            got = sub.f[:, :, gw:gw + (x1 - x0)]
            assert np.isfinite(got).all(), (k, "a value outside the exchanged set was needed")
            new.append(got.copy())
        slabs = new
        assert np.array_equal(np.concatenate(slabs, axis=2), whole.f)


@pytest.mark.parametrize("depth", [1, 2, 3])
@pytest.mark.parametrize("parts,dtype", [(2, np.float32), (3, np.float64)])
def test_published_columns_are_sufficient_for_a_launch_k_updates_deep(parts, depth, dtype):
    """The halo format as shipped (lb_fused.cuh, StepParams): per interior face a slab publishes all nine populations of
    its three outermost columns and the mask of its two outermost ones.  Model of a launch `depth` updates deep: every slab
    is advanced `depth` oracle steps on its own columns extended by `depth` ghost columns (NaN in the ghost columns a
    launch of that depth does not read, mask of depth-1 of them); its own columns come out NaN-free and bit-identical to
    the undivided lattice -- launches of mixed depth in one run, as lb_step plans them."""
    from oracle import oracle as orc
    from util import pipe_case
    orc.build()
    nx, ny = 90, 31
    f0, mask = pipe_case(orc, nx, ny, dtype, mask="touching", seed=8)
    mask[ny // 2 - 4: ny // 2 + 3, nx // parts - 3: nx // parts + 3] = 1             # a body across the first cut
    kw = dict(mask=mask, dtype=dtype)
    whole = orc.OpenCLSchemeOracle(f0, 1.3, 1.01, 1.0, **kw)
    cuts = [round(k * nx / parts) for k in range(parts + 1)]
    slabs = [f0[:, :, cuts[k]:cuts[k + 1]].copy() for k in range(parts)]
    for d in (depth, 1, depth, depth):                                                # e.g. 3 + 1 + 3 + 3 updates
        whole.run(d)
        published = [(s[:, :, :3].copy(), s[:, :, -3:].copy()) for s in slabs]       # (west three, east three) of each slab
        new = []
        for k, own in enumerate(slabs):
            x0, x1 = cuts[k], cuts[k + 1]
            gw = d if k > 0 else 0
            ge = d if k < parts - 1 else 0
            ext = np.full((9, ny, gw + (x1 - x0) + ge), np.nan, dtype)
            ext[:, :, gw:gw + (x1 - x0)] = own
            if gw:
                ext[:, :, :gw] = published[k - 1][1][:, :, 3 - gw:]                   # the west neighbour's east columns
            if ge:
                ext[:, :, gw + (x1 - x0):] = published[k + 1][0][:, :, :ge]
            m = np.zeros((ny, ext.shape[2]), np.uint8)
            mw, me = max(gw - 1, 0), max(ge - 1, 0)                                   # mask columns a launch of this depth reads
            assert mw <= 2 and me <= 2
            m[:, gw - mw: gw + (x1 - x0) + me] = mask[:, x0 - mw: x1 + me]
            sub = orc.OpenCLSchemeOracle(ext, 1.3, 1.01, 1.0, mask=m, dtype=dtype)
            with np.errstate(all="ignore"):
                sub.run(d)
            got = sub.f[:, :, gw:gw + (x1 - x0)]
            assert np.isfinite(got).all(), (k, d)
            new.append(got.copy())
        slabs = new
        assert np.array_equal(np.concatenate(slabs, axis=2), whole.f), d


def test_marching_launch_geometry_covers_every_row_once():
    """lb_plan_march_launch = the launchers' own arithmetic (csrc/lb_host.h), no device needed: tall segments followed by
    short ones cover the row range exactly once whatever the sizes; small ranges stay uniform; a last strip narrower than
    the three published columns makes two east edge strips."""
    import ctypes as ct
    from lb_b200 import native
    L = native.lib()
    rng = np.random.RandomState(12)

    def plan(nx, rows, elem=4, depth=3, nw=4, minb=4, s1=64, s2=16, sms=148, w=0, e=0):
        out = [ct.c_int() for _ in range(5)]
        rc = L.lb_plan_march_launch(nx, rows, elem, depth, nw, minb, s1, s2, sms, w, e, *[ct.byref(o) for o in out])
        assert rc == 0
        return [o.value for o in out]

    graded = 0
    for _ in range(400):
        nx, rows = int(rng.randint(1, 40000)), int(rng.randint(1, 70000))
        s1 = int(rng.choice([8, 11, 16, 32, 64, 128])); s2 = int(rng.choice([0, s1 // 4, s1 // 2, s1]))
        elem, depth = int(rng.choice([4, 8])), int(rng.choice([2, 3]))
        nstrips, ne, n_tall, n_short, short = plan(nx, rows, elem, depth, 4, int(rng.choice([4, 5, 6])), s1, s2)
        out = 120 if elem == 4 else (56 if depth == 3 else 60)
        assert nstrips == -(-nx // out) and ne == 0
        # the kernels' mapping from segment index to rows (lb_march.cuh)
        covered = np.zeros(rows, dtype=np.int32)
        for seg in range(n_tall + n_short):
            tall = seg < n_tall
            ys = seg * s1 if tall else n_tall * s1 + (seg - n_tall) * short
            ye = min(ys + (s1 if tall else short), rows)
            assert ys < rows, (nx, rows, s1, s2, seg)
            covered[ys:ye] += 1
        assert (covered == 1).all(), (nx, rows, s1, s2)
        if n_short:
            graded += 1
            assert short == s2 < s1 and n_short * short < rows // 2 + s1 + short
        else:
            assert short == s1 and n_tall == -(-rows // s1)
    assert graded > 20
    # C4 on one GPU, its N=8 slabs, a lattice too small for short segments
    assert plan(32768, 32768, s1=128, s2=32)[2:] == [251, 20, 32]
    assert plan(4096, 32768, s1=64, s2=16)[2:] == [478, 136, 16]
    assert plan(4096, 1024, depth=2, minb=6, s1=11, s2=0)[2:] == [94, 0, 11]
    # edge strips: none on a single slab; west + east; a last strip of one or two columns adds the one before it
    assert plan(4096, 100, w=1, e=1)[1] == 2 and plan(4096, 100, w=0, e=1)[1] == 1 and plan(4096, 100, w=1, e=0)[1] == 1
    for nx, want in ((121, 2), (122, 2), (123, 1), (240, 1), (241, 2), (120, 1), (2, 1)):
        assert plan(nx, 100, w=0, e=1)[1] == want, nx
    assert plan(121, 100, w=1, e=1)[1] == 2 and plan(241, 100, w=1, e=1)[1] == 3
    assert plan(57, 100, elem=8, depth=3, w=0, e=1)[1] == 2 and plan(59, 100, elem=8, depth=3, w=0, e=1)[1] == 1
    assert plan(61, 100, elem=8, depth=2, w=0, e=1)[1] == 2
