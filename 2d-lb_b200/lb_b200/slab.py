"""x-slab decomposition across GPUs, one process per GPU (`torch.distributed` for the plumbing).

The reference is single-device (no communication layer at all); this is the
multi-GPU design `north_star` asks for.  Rank r owns global columns [x_offset, x_offset + nx) for
all y and all nine populations.  Per step and per interior face three populations cross:
1,5,8 eastward and 3,6,7 westward, 3*ny*sizeof(T) bytes each way.

Data path: none of it goes through a collective.  At construction the ranks exchange the CUDA IPC
handles of their halo arenas (all_gather_object); afterwards the fused kernel's boundary threads
store their outgoing populations straight into the neighbour's ghost column over NVLink and
publish a step flag there, and the neighbour's boundary tiles poll that flag locally before
reading (lb_fused.cuh).  `torch.distributed` is used for the rendezvous, barriers and for
gathering results in tests -- never inside the step loop.
"""
import numpy as np

from . import native as N
from .lattice import Lattice, slab_edges, split_slabs


class SlabLattice:
    """This rank's slab of a global (global_nx x ny) lattice.

    dist             an initialised torch.distributed module/process group wrapper exposing
                     get_rank / get_world_size / all_gather_object / barrier (default: torch.distributed)
    lattice_factory  callable creating the per-rank lattice (default `Lattice`; tests inject a fake
                     to exercise the host logic under gloo without a GPU)
    """

    def __init__(self, global_nx, ny, omega, inlet_rho=1.0, outlet_rho=1.0, bc="pipe", dtype=np.float32,
                 math="strict", device=None, zero_obstacle_velocity=False, stream=None, dist=None,
                 lattice_factory=Lattice):
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.global_nx, self.ny, self.bc = int(global_nx), int(ny), bc
        self.ranges = split_slabs(global_nx, self.world)
        self.x_offset, self.nx = self.ranges[self.rank]
        self.west_edge, self.east_edge = slab_edges(self.rank, self.world, bc)
        self.device = self.rank if device is None else device
        self.lat = lattice_factory(self.nx, ny, omega, inlet_rho, outlet_rho, bc=bc, dtype=dtype, math=math,
                                   device=self.device, zero_obstacle_velocity=zero_obstacle_velocity,
                                   global_nx=global_nx, x_offset=self.x_offset, west_edge=self.west_edge,
                                   east_edge=self.east_edge, stream=stream)
        self._primed = False
        self._connect()

    # -- rendezvous ---------------------------------------------------------------------
    def neighbours(self):
        """(west_rank, east_rank); None where the slab touches the domain boundary."""
        w = (self.rank - 1) % self.world if self.west_edge == "halo" else None
        e = (self.rank + 1) % self.world if self.east_edge == "halo" else None
        return w, e

    def _connect(self):
        if self.world == 1:
            return
        mine = (self.lat.halo_ipc_handle(), self.device)
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        w, e = self.neighbours()
        if w is not None:
            self.lat.halo_connect_ipc("west", everyone[w][0], everyone[w][1])
        if e is not None:
            self.lat.halo_connect_ipc("east", everyone[e][0], everyone[e][1])
        self.dist.barrier()

    def prime(self):
        """Publish the current boundary columns to the neighbours (after any upload/initialiser)."""
        self._primed = True
        if self.world == 1:
            return
        self.lat.sync()
        self.dist.barrier()          # every rank's state is final before anyone writes ghosts
        self.lat.halo_prime()
        self.dist.barrier()          # every ghost column is filled before anyone steps

    # -- slab-local views of global host arrays -----------------------------------------------
    def local(self, a):
        """Slice the trailing (x) axis of a global device-layout array down to this slab."""
        return np.ascontiguousarray(a[..., self.x_offset:self.x_offset + self.nx])

    def set_mask(self, global_mask):
        self.lat.set_mask(self.local(np.asarray(global_mask)))
        self._primed = False         # the neighbours hold a copy of this slab's boundary mask column

    def set_mask_disk(self, cx, cy, r):
        self.lat.set_mask_disk(cx, cy, r)
        self._primed = False

    def upload_f(self, global_f):
        self.lat.upload_f(self.local(np.asarray(global_f)))
        self.prime()

    # -- hot path ---------------------------------------------------------------------------
    def run(self, n, sync=True):
        if not self._primed:
            self.prime()             # collective, like every call that precedes it on all ranks
        self.lat.run(n, sync=sync)

    def sync(self):
        self.lat.sync()

    def gather(self, field):
        """Global array on every rank (test/diagnostic helper, not for production-size grids)."""
        part = self.lat.download(field)
        parts = [None] * self.world
        self.dist.all_gather_object(parts, part)
        return np.concatenate(parts, axis=-1)

    def total_mass(self):
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.lat.total_mass())
        return float(sum(parts))

    def checksum(self):
        """Exact checksum of the whole lattice: the per-slab 64-bit sums, added modulo 2^64."""
        parts = [None] * self.world
        self.dist.all_gather_object(parts, self.lat.checksum())
        return sum(parts) & 0xFFFFFFFFFFFFFFFF

    def close(self):
        """Collective: every rank drains its stream, then all ranks meet, then the arenas go away -- a
        neighbour's last launch stores into this rank's arena until that neighbour has synchronised."""
        if self.lat is None:
            return
        try:
            self.lat.sync()
        finally:
            if self.world > 1:
                self.dist.barrier()
            self.lat.close()
            self.lat = None


__all__ = ["SlabLattice", "N"]
