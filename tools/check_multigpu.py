#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/check_multigpu.py

For a pipe flow with obstacles and a periodic box (fp32 and fp64) the x-slab run over N GPUs with
peer-memory halos -- one-update kernel and two-update marching kernel -- must be BIT-IDENTICAL to the
single-slab one-update run of the same global lattice (rank 0 computes that one on its own GPU).
Exit code 0 = all identical.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "2d-lb_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lb_b200 import Lattice  # noqa: E402
from lb_b200.slab import SlabLattice  # noqa: E402


def make_case(bc, dtype, nx, ny, seed):
    rng = np.random.RandomState(seed)
    w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
    if bc == "pipe":
        rho = (1.01 - np.arange(nx) * 0.01 / nx)[None, :].repeat(ny, 0)
        mask = (rng.rand(ny, nx) < 0.03).astype(np.uint8)
        mask[:, :2] = 0
        mask[:, -2:] = 0
    else:
        rho = np.ones((ny, nx))
        mask = None
    f0 = (w[:, None, None] * rho[None] * (1 + 1e-3 * rng.randn(9, ny, nx))).astype(dtype)
    return f0, mask


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for bc in ("pipe", "periodic"):
        for dtype in (np.float32, np.float64):
            for math, tb in (("strict", "off"), ("strict", "march.w4b5.sh.s64"), ("fast", "march.w4b4.s32"), ("fast", "off"),
                             ("strict", "march3.w4b5.s16")):
                # 121-122 columns per slab: the published columns spread over the last TWO strips of a slab
                for nx in (64 * world + 37, 121 * world + world // 2):
                    ny, steps = 301, 61
                    f0, mask = make_case(bc, dtype, nx, ny, seed=11)
                    slab = SlabLattice(nx, ny, 1.5, 1.01, 1.0, bc=bc, dtype=dtype, math=math, device=local)
                    slab.lat.set_temporal_blocking(tb)
                    if mask is not None:
                        slab.set_mask(mask)
                    slab.upload_f(f0)
                    slab.run(steps)
                    slab.run(steps - 1)
                    got = {k: slab.gather(k) for k in ("f", "rho", "u")}
                    mass = slab.total_mass()
                    slab.close()
                    if rank == 0:
                        with Lattice(nx, ny, 1.5, 1.01, 1.0, mask=mask, f0=f0, bc=bc, dtype=dtype, math=math, device=local) as one:
                            one.set_temporal_blocking("off")
                            one.run(2 * steps - 1)
                            same = all(np.array_equal(got[k], one.download(k)) for k in got)
                            m1 = one.total_mass()
                        print(f"[check_multigpu] N={world} {bc:8s} {nx}x{ny} {np.dtype(dtype).name} {math:6s} {tb:15s}: "
                              f"{'bit-identical' if same else 'MISMATCH'}  mass {mass:.10e} vs {m1:.10e}", flush=True)
                        ok = ok and same
                    dist.barrier()
    # single-process multi-device path behind the drop-in classes (rank 0 drives all visible GPUs)
    if rank == 0 and torch.cuda.device_count() >= 2:
        import lb_b200.dimensionless as lb
        devs = list(range(min(torch.cuda.device_count(), 4)))
        kw = dict(cylinder_center=[0.75, 0.5], cylinder_radius=0.1, diameter=1., rho=1., viscosity=1., pressure_grad=-10.,
                  pipe_length=3., N=20, time_prefactor=4., verbose=False)
        np.random.seed(3)
        single = lb.Pipe_Flow_Cylinder(device=local, **kw)
        single.run(151)
        fs = single.get_fields()
        for tb in ("off", "march.w4b5.sh.s32"):
            np.random.seed(3)
            multi = lb.Pipe_Flow_Cylinder(devices=devs, **kw)        # lb_multi_* underneath: one slab per device
            multi.sim.set_temporal_blocking(tb)
            multi.run(151)
            fm = multi.get_fields()
            same = all(np.array_equal(fm[k], fs[k]) for k in ("f", "feq", "rho", "u", "v"))
            print(f"[check_multigpu] single-process Pipe_Flow_Cylinder(devices={devs}) {multi.nx}x{multi.ny} {tb}: "
                  f"{'bit-identical' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
            multi.sim.close()
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("[check_multigpu]", "PASS" if flag.item() else "FAIL")
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
