#!/usr/bin/env python
"""Benchmark of the D2Q9 hot path (one fused collide-and-stream kernel per lattice update -- per TWO updates on
lattices of at least 2^22 nodes, where lb_step runs the marching kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c3|c5] [--impl reference]

For N > 1 launch as the driver does:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Prints ONE JSON line (rank 0).  Metric: MLUPS = lattice updates / second / 1e6, the reference's own
definition (docs/python_cython_opencl_comparison.ipynb cell 16).  A "step" is one lattice update of
the whole grid.  Workloads are BASELINE.json's configs:
  c4 (default) cylinder wake 32768x32768 fp32, x-slab STRONG scaling over the N GPUs
  c1           Poiseuille 256x128 fp64 (the reference's CPU-runnable case; launch-bound on a GPU)
  c2           Pipe_Flow_Obstacles-style 4096x1024 fp32 with an obstacle mask (1 GPU)
  c3           periodic shear layers (Kelvin-Helmholtz) 16384x16384 fp32 (1 GPU)
  c5           channel flow 16384x16384 fp64 PER GPU, weak scaling
  pub          the reference's published benchmark through the drop-in class API (vs_baseline = / 317.5 MLUPS)
`value`  : device-resident throughput, CUDA events around K launches, max over ranks.
`e2e`    : the same K steps through the C-ABI with HOST buffers: upload of f from pinned memory, K steps,
           read-back of rho, u, v -- all inside the timed region (N = 1: one pipelined lb_run_streamed call;
           N > 1: lb_upload_f + lb_halo_prime + lb_step + lb_download per rank); `equals_resident_run`: the
           populations the pipelined call leaves behind have the checksum of upload + run (checked untimed, N = 1).
`roofline`: achieved = ALGORITHMIC bytes per launch (72 B fp32 / 144 B fp64 per lattice update x updates per
           launch) over the average launch duration, against MEASURED_PEAKS.json's hbm_gbs; `traffic` = the
           kernel's measured DRAM bytes per launch (ncu, profiles/traffic.json) and `frac_on_measured_traffic`
           what that amounts to -- the two-update kernel moves about half the algorithmic bytes.
`checks.checksum`: exact 64-bit checksum of the final populations, summed over ranks: equal at every N.
`cpu_baseline`: the unmodified reference Cython path (oracle/_ref) on one host core, bounded sample.
`--impl reference`: the reference's CPU path on all host cores (independent replicas; the
           reference has no threaded path), same metric/config keys.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "2d-lb_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (description, global nx per GPU count fn, ny, dtype, bc, scaling, omega, inlet_rho, init, mask)
    "c4": dict(desc="cylinder wake 32768x32768 fp32, x-slab strong scaling", nx=32768, ny=32768, dtype="f32",
               bc="pipe", scaling="strong", omega=1.7, inlet_rho=1.003, init="pipe_ramp", mask="disk"),
    "c1": dict(desc="Poiseuille pipe flow 256x128 fp64 (launch-bound: 32-step CUDA graphs)", nx=256, ny=128, dtype="f64",
               bc="pipe", scaling="strong", omega=1.000265, inlet_rho=1.00495022, init="pipe_ramp", mask=None),
    "c2": dict(desc="Pipe_Flow_Obstacles 4096x1024 fp32 with obstacle mask", nx=4096, ny=1024, dtype="f32",
               bc="pipe", scaling="strong", omega=1.0, inlet_rho=1.01, init="pipe_ramp", mask="cs205", zero_vel=True),
    "c3": dict(desc="periodic vortex-sheet (Kelvin-Helmholtz) 16384x16384 fp32", nx=16384, ny=16384, dtype="f32",
               bc="periodic", scaling="strong", omega=1.7, inlet_rho=1.0, init="shear_layers", mask=None),
    "c5": dict(desc="weak-scaling channel flow 16384x16384 per GPU, fp64", nx=16384, ny=16384, dtype="f64",
               bc="pipe", scaling="weak", omega=1.0, inlet_rho=1.001, init="pipe_ramp", mask=None),
}
PUBLISHED_REFERENCE_MLUPS = 317.5   # OpenCL path on a GTX Titan Black, 3751x1251 (other hardware/config)


def run_published_config(args):
    """`--workload pub`: the reference's OWN published benchmark, unchanged user code
    (docs/python_cython_opencl_comparison.ipynb cells 10 and 16): Pipe_Flow_Cylinder, N=125 ->
    3751x1251 fp32, `sim.run(num_steps)` timed with the wall clock, MLUPS as the notebook computes
    it.  Published: 317.5 MLUPS on a GTX Titan Black -> vs_baseline."""
    import numpy as np
    import torch
    from LB_D2Q9.dimensionless import opencl_dim as lb_cl
    torch.cuda.set_device(0)
    N = 125
    D, rho, nu, pressure_grad = 1., 1., 1., -10
    pipe_length = 3 * D
    cylinder_center = [pipe_length / 4, D / 2]
    cylinder_radius = D / 10
    np.random.seed(0)
    t_ctor = time.time()
    sim_cl = lb_cl.Pipe_Flow_Cylinder(diameter=D, rho=rho, viscosity=nu, pressure_grad=pressure_grad, pipe_length=pipe_length,
                                      N=N, time_prefactor=1., cylinder_center=cylinder_center, cylinder_radius=cylinder_radius,
                                      two_d_local_size=(32, 32), three_d_local_size=(32, 32, 1), verbose=False)
    t_ctor = time.time() - t_ctor
    total_lattice_size = sim_cl.nx * sim_cl.ny
    num_steps = max(args.steps, 1000)
    sim_cl.run(max(args.warmup, 3))
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.2)
    launches0 = sim_cl.sim.launch_count
    start_time = time.time()
    sim_cl.run(num_steps)                       # one host sync at the end, like the notebook's blocking run
    time_elapsed = time.time() - start_time
    launches = sim_cl.sim.launch_count - launches0
    # a longer region for the clock sampler (the timed run above lasts ~60 ms)
    sim_cl.run(20 * num_steps)
    clocks = sampler.stop()
    mlups = (num_steps * (total_lattice_size / 10. ** 6.)) / time_elapsed
    t0 = time.time()
    fields = sim_cl.get_fields()                # f, feq, u, v, rho to the host, like the notebook's readback
    t_read = time.time() - t0
    peak, peak_src = measured_peak()
    achieved = total_lattice_size * 72 / (time_elapsed / num_steps) / 1e9
    e2e_mlups = (num_steps * (total_lattice_size / 10. ** 6.)) / (t_ctor + time_elapsed + t_read)
    line = {
        "metric": "D2Q9 MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": 1, "steps": num_steps, "warmup": max(args.warmup, 3),
        "ms_per_step": time_elapsed * 1e3 / num_steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": mlups / PUBLISHED_REFERENCE_MLUPS, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pub: the reference's published benchmark, Pipe_Flow_Cylinder N=125 -> 3751x1251 fp32, "
                               "sim.run(1000) by wall clock through the drop-in class API",
                   "grid": [sim_cl.nx, sim_cl.ny], "omega": float(sim_cl.omega), "math": "strict",
                   "baseline": "317.5 MLUPS, GTX Titan Black, docs/python_cython_opencl_comparison.ipynb cell 16",
                   "l2": "working set 2 x 169 MB ping-pong > 126 MB L2 (partly L2-resident: read `roofline` accordingly)"},
        "clocks": clocks,
        "e2e": {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": 9 * total_lattice_size * 4 / num_steps,
                "d2h_bytes_per_step": 21 * total_lattice_size * 4 / num_steps,
                "call": "constructor (host init + upload) + run(%d) + get_fields() (f, feq, u, v, rho)" % num_steps,
                "constructor_s": t_ctor, "readback_s": t_read},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0},
        "cpu_baseline": None,
        "checks": {"rho_finite": bool(np.isfinite(fields["rho"]).all())},
    }
    emit(line)
    return 0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (torch copy_, burst)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="lb_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, val in zip(names, c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
# reference CPU path (oracle/_ref = the unmodified compiled reference; test infrastructure, used
# here only as the timed CPU arm)
# ---------------------------------------------------------------------------------------------
CPU_BASELINE_STEPS = 600
CPU_SAMPLE = "reference cython_dim.Pipe_Flow_Cylinder 751x251 (the cylinder-wake workload at N=25), fp32 storage"


def _ref_sim():
    import numpy as np
    from oracle import refload
    cd = refload.cython_dim()
    np.random.seed(0)
    with refload.quiet():
        sim = cd.Pipe_Flow_Cylinder(cylinder_center=[0.75, 0.5], cylinder_radius=0.1, diameter=1., rho=1., viscosity=1.,
                                    pressure_grad=-10., pipe_length=3., N=25)
    return sim


def _port_sim():
    """Fallback when oracle/_ref is not available: the C restatement (kind 'port')."""
    import numpy as np
    from oracle import oracle as orc
    nx, ny = 751, 251
    rho = np.ones((ny, nx))
    f0 = orc.feq_of(rho, np.zeros((ny, nx)), np.zeros((ny, nx)), np.float32)
    mask = np.zeros((ny, nx), np.int32)
    yy, xx = np.ogrid[0:ny, 0:nx]
    mask[(xx - nx / 4) ** 2 + (yy - ny / 2) ** 2 < (ny / 10) ** 2] = 1
    sim = orc.OpenCLSchemeOracle(f0, 0.4132, 1.0037, 1.0, mask=mask)
    sim.nx, sim.ny = nx, ny
    return sim


def _cpu_worker(steps, warmup, conn):
    try:
        from oracle import refload
        kind = "reference" if refload.available() else "port"
        sim = _ref_sim() if kind == "reference" else _port_sim()
        sim.run(warmup)
        conn.send(("ready", kind, sim.nx * sim.ny))
        conn.recv()                       # start signal
        t0 = time.perf_counter()
        sim.run(steps)
        conn.send(("done", time.perf_counter() - t0))
    except Exception as exc:              # pragma: no cover
        conn.send(("error", repr(exc)))


def time_reference_cpu(steps, warmup, replicas):
    """`replicas` independent simulations, one process each, started together.  Returns
    (aggregate MLUPS, seconds, kind, cells per replica)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for _ in range(replicas):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_cpu_worker, args=(steps, warmup, b))
        pr.start()
        pipes.append(a)
        procs.append(pr)
    kind, cells = None, None
    for a in pipes:
        msg = a.recv()
        if msg[0] != "ready":
            raise RuntimeError(f"CPU reference worker failed: {msg}")
        kind, cells = msg[1], msg[2]
    t0 = time.perf_counter()
    for a in pipes:
        a.send("go")
    for a in pipes:
        msg = a.recv()
        if msg[0] != "done":
            raise RuntimeError(f"CPU reference worker failed: {msg}")
    wall = time.perf_counter() - t0
    for pr in procs:
        pr.join(timeout=30)
    return cells * steps * replicas / wall / 1e6, wall, kind, cells


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warmup = args.steps, max(args.warmup, 1)
    mlups, wall, kind, cells = time_reference_cpu(steps, warmup, cores)
    line = {
        "impl": "reference", "metric": "D2Q9 MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": wall * 1e3 / steps, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}",
                   "note": "CPU arm runs a bounded sample of the workload; MLUPS is size-normalised"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind,
                         "sample": f"{cores} independent replicas (the reference has no threaded path) of {CPU_SAMPLE}, "
                                   f"{steps} steps each"},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# the CUDA arm
# ---------------------------------------------------------------------------------------------
_RESULT_FD = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner on fd 1 when NCCL_DEBUG is set in the environment), so everything that is not the result goes
    to stderr: fd 1 is pointed at fd 2 and the original stdout is kept for emit()."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    payload = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, payload)
    else:
        os.write(_RESULT_FD, payload)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + ["pub"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="strict", choices=["fast", "strict"])
    ap.add_argument("--variant", default="")
    ap.add_argument("--tb2", default="auto", help="temporal blocking tile: auto (default), off, or a tile name (rows14.w8 ...)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nx", type=int, default=0, help="override the workload's global nx (debug)")
    ap.add_argument("--ny", type=int, default=0)
    args = ap.parse_args()
    if args.workload == "pub":
        if args.impl == "reference":
            args.workload = "c4"
        else:
            return run_published_config(args)
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return 0
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks", file=sys.stderr)
            return 2
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    from lb_b200 import native
    from lb_b200.slab import SlabLattice

    if native.lib().lb_device_count() == 0:
        print("bench.py: no CUDA device; the product path has no CPU fallback", file=sys.stderr)
        return 3
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    class _Solo:                      # torch.distributed stand-in for a single process
        @staticmethod
        def get_rank():
            return 0

        @staticmethod
        def get_world_size():
            return 1

    gny = args.ny or wl["ny"]
    gnx = args.nx or wl["nx"]
    if wl["scaling"] == "weak":
        gnx *= world
    dtype = np.float32 if wl["dtype"] == "f32" else np.float64
    elem = 4 if wl["dtype"] == "f32" else 8
    bytes_per_lu = 18 * elem
    stream = torch.cuda.Stream()
    slab = SlabLattice(gnx, gny, wl["omega"], wl["inlet_rho"], 1.0, bc=wl["bc"], dtype=dtype, math=args.math,
                       device=local_rank, zero_obstacle_velocity=bool(wl.get("zero_vel")), stream=stream.cuda_stream,
                       dist=dist if world > 1 else _Solo)
    lat = slab.lat
    if args.variant:
        lat.set_variant(args.variant)
    if args.tb2 != "auto":
        lat.set_temporal_blocking(args.tb2)           # same choice on every rank (the slabs exchange per LAUNCH)
    if wl["mask"] == "disk":
        lat.set_mask_disk(gnx / 4.0, gny / 2.0, gny / 10.0)
    elif wl["mask"] == "cs205":
        # docs/cs205_binary.tif of the reference (bit-packed fixture), nearest-neighbour resampled
        # to the lattice: mask[x, y] = src[x*800//nx, y*400//ny]   
        from lb_b200 import masks
        src = masks.unpack(np.load(os.path.join(ROOT, "tests", "golden", "cs205_binary_mask.npz")))
        slab.set_mask(np.ascontiguousarray(masks.resample(src, gnx, gny).T))
    lat.init_synthetic(wl["init"], u0=0.05, amplitude=1e-3, seed=2015)
    slab.prime()
    cells_global = gnx * gny
    cells_local = slab.nx * gny

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """barrier+sync | events on the launch stream | barrier+sync ; max over ranks, in ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            e0.record()
            fn()
            e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident throughput -----------------------------------------------------------
    lat.run(args.warmup)
    launches0 = lat.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ms = timed(lambda: lat.run(args.steps, sync=False))
    clocks = sampler.stop() if rank == 0 else None
    launches = lat.launch_count - launches0
    lat.sync()
    mass = slab.total_mass() if world > 1 else lat.total_mass()
    checksum = slab.checksum() if world > 1 else lat.checksum()
    value = cells_global * args.steps / (ms * 1e-3) / 1e6
    launch_ms = ms / args.steps
    achieved = cells_local * bytes_per_lu / (launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- what an arithmetic-free kernel with the same access pattern reaches, here and now ---------
    ceiling = None
    try:
        barrier()
        c_ms = lat.copy_ceiling_ms(10)
        if world > 1:
            t = torch.tensor([c_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c_ms = float(t.item())
        ceiling = {"GB/s": cells_local * bytes_per_lu / (c_ms * 1e-3) / 1e9, "ms_per_launch": c_ms,
                   "how": "lb_selftest_copy: 9 x 128-bit loads (D2Q9 row offsets) + 9 x 128-bit stores per thread, "
                          "the fused kernel's thread mapping, no arithmetic; same buffers, 10 launches, CUDA events"}
    except Exception as exc:                      # a diagnostic: never fails the bench
        ceiling = {"GB/s": None, "error": repr(exc)}

    # ---- end to end through the C-ABI with host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        n_f, n_m = 9 * cells_local, cells_local
        tdt = torch.float32 if elem == 4 else torch.float64
        host_f = torch.empty(n_f, dtype=tdt, pin_memory=True)
        host_m = [torch.empty(n_m, dtype=tdt, pin_memory=True) for _ in range(3)]
        f_np = host_f.numpy().reshape(9, gny, slab.nx)
        m_np = [h.numpy().reshape(gny, slab.nx) for h in host_m]
        lat.download("f", out=f_np)          # the step's input, resident in pinned host memory (untimed)

        def e2e_call():
            if world == 1:
                # one C-ABI call: upload, K steps, read-back, pipelined by row bands (lb_run_streamed)
                lat.run_streamed(f_np, args.steps, rho=m_np[0], u=m_np[1], v=m_np[2])
                return
            lat.upload_f(f_np)               # H2D of all nine populations (blocking)
            slab.prime()                     # republish the fresh state's boundary columns
            lat.run(args.steps, sync=False)
            for name, out in zip(("rho", "u", "v"), m_np):
                lat.download(name, out=out)  # D2H of density and velocity (blocking)

        e2e_call()                           # warm-up (page-locks, graph builds)
        ms_e2e = timed(e2e_call)
        ok = bool(np.isfinite(m_np[0]).all())
        same = None
        if world == 1:                       # the pipelined call against upload + run, untimed: the same bits?
            c_streamed = lat.checksum()
            lat.upload_f(f_np)
            lat.run(args.steps)
            same = c_streamed == lat.checksum()
        e2e = {"value": cells_global * args.steps / (ms_e2e * 1e-3) / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": n_f * elem * world / args.steps, "d2h_bytes_per_step": 3 * n_m * elem * world / args.steps,
               "ms_per_call": ms_e2e, "steps_per_call": args.steps, "rho_finite": ok, "equals_resident_run": same,
               "call": ("lb_run_streamed(pinned host f, K, pinned rho, u, v): upload, K steps and read-back pipelined by row bands"
                        if world == 1 else "lb_upload_f(pinned host f) + lb_halo_prime + lb_step(K) + lb_download(rho,u,v) per call")}
        del host_f, host_m

    # ---- CPU baseline (rank 0, N == 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, wall, kind, cells = time_reference_cpu(CPU_BASELINE_STEPS, 2, 1)
            cpu = {"value": v, "unit": "MLUPS", "cores": 1, "kind": kind,
                   "sample": f"{CPU_SAMPLE}, {CPU_BASELINE_STEPS} steps, 1 thread ({wall:.1f} s) of {os.cpu_count()} host cores available"}
        except Exception as exc:
            cpu = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "unavailable", "sample": repr(exc)}

    if rank == 0:
        tb2 = lat.temporal_blocking
        two = tb2 != "off"
        updates_per_launch = 1 if not two else (3 if tb2.startswith("march3") else 2)
        kernel = ("fused_step_kernel (one lattice update per launch)" if not two else
                  f"fused_march{'_k' if updates_per_launch == 3 else ''}_kernel shape {tb2}, {lat.segment_rows}-row segments ({'three' if updates_per_launch == 3 else 'two'} "
                  "lattice updates per launch, the intermediate time levels on chip; a run whose length is not a multiple of that "
                  "starts with one shorter launch)")
        # measured DRAM traffic of ONE launch of the dominant kernel (ncu --set full, dram__bytes_read.sum +
        # dram__bytes_write.sum), recorded per lattice node in profiles/traffic.json with the capture's own grid
        traffic, traffic_note = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                t = json.load(fh).get(wl["dtype"] + ("" if not two else ("_march3" if updates_per_launch == 3 else "_march")))
            if t:
                traffic = t["dram_bytes_per_node_per_launch"] * cells_local
                same = list(t.get("grid", [])) == [slab.nx, gny]
                traffic_note = ("ncu capture of this kernel on this grid: " if same else
                                f"ncu capture of this kernel on {t.get('grid')}, scaled by node count: ") + t["source"]
        except Exception:
            pass
        launch_dur_ms = launch_ms * updates_per_launch
        line = {
            "metric": "D2Q9 MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "global_grid": [gnx, gny],
                       "per_gpu_grid": [slab.nx, gny], "bc": wl["bc"], "omega": wl["omega"], "math": args.math,
                       "decomposition": f"x-slabs x{world}, peer-memory halos over NVLink (27 values per face row per launch)",
                       "kernel": kernel,
                       "l2": f"inputs larger than L2: {2 * 9 * cells_local * elem / 1e9:.1f} GB ping-pong working set per GPU vs 126 MB",
                       "published_reference_mlups_other_hw": PUBLISHED_REFERENCE_MLUPS},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note,
                         "dram_GBs_on_measured_traffic": (traffic / (launch_dur_ms * 1e-3) / 1e9) if traffic else None,
                         "frac_on_measured_traffic": (traffic / (launch_dur_ms * 1e-3) / 1e9 / peak) if traffic else None,
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "bytes_per_lattice_update": bytes_per_lu, "updates_per_launch": updates_per_launch,
                         "launch_ms": launch_dur_ms, "algorithmic_bytes_per_launch": cells_local * bytes_per_lu * updates_per_launch,
                         "per": "GPU, " + kernel.split(" ")[0],
                         "note": None if not two else "achieved = ALGORITHMIC bytes (9 loads + 9 stores per update x updates per "
                                 "launch) / launch duration; the kernel keeps the intermediate levels on chip and moves a half "
                                 "or a third of that through HBM (traffic), so frac can exceed 1 -- frac_on_measured_traffic is "
                                 "the share of the measured HBM peak the kernel's real DRAM traffic amounts to",
                         "pattern_copy_ceiling": ceiling,
                         "frac_of_pattern_copy_ceiling": (achieved / ceiling["GB/s"]) if ceiling and ceiling.get("GB/s") else None},
            "cpu_baseline": cpu,
            "checks": {"total_mass": mass, "mass_finite": bool(np.isfinite(mass)),
                       "checksum": f"{checksum:016x}",
                       "checksum_of": f"populations after {args.warmup} + {args.steps} steps from the seeded device-side initial "
                                      "state: 64-bit wrap-around sum of all bit patterns, summed over ranks -- equal values at "
                                      "every N (and with --tb2 off) mean bit-identical results"},
        }
        emit(line)
    slab.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
