"""Multi-GPU slab decomposition of the physics step, host side.

Rank g owns grid rows [g*ny/G, (g+1)*ny/G).  libplife.so packs / unpacks the halo and migration
messages (csrc/slab.cu); this module moves them between ranks:

  * `DistExchange`   one process per GPU, torch.distributed send/recv (NCCL over NVLink on GPUs,
                     gloo on CPU for the host-logic tests);
  * `LocalExchange`  several virtual ranks inside one process (device-to-device copies), which
                     lets a single GPU exercise the whole slab path against the single-GPU result.

Message routing (down = rank-1, up = rank+1, periodic when wrap is on):
  send[0] -> down neighbour's recv[1]      send[1] -> up neighbour's recv[0]
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N


# ---------------------------------------------------------------------------
# ownership (must match csrc/plife_api.cu:make_grid and cells.cu:container_of)
# ---------------------------------------------------------------------------

def grid_rows(rmax: float) -> int:
    """nx = ny = floor(1/rmax) (B/Physics.java:82-85)."""
    return int(np.floor(1.0 / rmax))


def slab_rows(rank: int, world: int, ny: int):
    return rank * ny // world, (rank + 1) * ny // world


def owner_of_row(cy, world: int, ny: int):
    """Rank whose slab contains global row cy: the largest g with g*ny//world <= cy."""
    return ((np.asarray(cy, np.int64) + 1) * world - 1) // ny


def owner_of_position(y, rmax: float, world: int, fp32: bool = True):
    """Owner rank of particles by their y coordinate (fp32 storage rounds first, like the upload)."""
    y = np.asarray(y, np.float64)
    if fp32:
        y = y.astype(np.float32).astype(np.float64)
    ny = grid_rows(rmax)
    cy = np.minimum((y / rmax).astype(np.int64), ny - 1)  # B/Physics.java:364,370-372
    return owner_of_row(cy, world, ny)


def neighbours(rank: int, world: int, wrap: bool):
    """(down, up) ranks or None across a closed boundary / for a single rank."""
    if world <= 1:
        return None, None
    dn = (rank - 1) % world if (wrap or rank > 0) else None
    up = (rank + 1) % world if (wrap or rank < world - 1) else None
    return dn, up


# ---------------------------------------------------------------------------
# exchanges
# ---------------------------------------------------------------------------

class DistExchange:
    """send[0] -> down.recv[1], send[1] -> up.recv[0] with torch.distributed P2P.

    The receive order (recv[1] first, then recv[0]) makes the two-rank case, where both
    neighbours are the same peer, match the peer's send order (send[0], then send[1])."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group

    def exchange(self, send: Sequence, recv: Sequence, wrap: bool, what: str = ""):
        import torch.distributed as dist
        dn, up = neighbours(self.rank, self.world, wrap)
        ops = []
        if dn is not None:
            ops.append(dist.P2POp(dist.isend, send[0], dn, self.group))
        if up is not None:
            ops.append(dist.P2POp(dist.isend, send[1], up, self.group))
        if up is not None:
            ops.append(dist.P2POp(dist.irecv, recv[1], up, self.group))
        if dn is not None:
            ops.append(dist.P2POp(dist.irecv, recv[0], dn, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class LocalExchange:
    """Exchange between virtual ranks living in this process (copies on the current stream)."""

    def __init__(self, world: int):
        self.world = world

    def exchange_all(self, sends: List[Sequence], recvs: List[Sequence], wrap: bool):
        for r in range(self.world):
            dn, up = neighbours(r, self.world, wrap)
            if dn is not None:
                recvs[dn][1].copy_(sends[r][0])
            if up is not None:
                recvs[up][0].copy_(sends[r][1])


# ---------------------------------------------------------------------------
# one slab
# ---------------------------------------------------------------------------

class SlabPhysics:
    """One rank of the slab-decomposed simulation: a libplife handle in slab mode.

    exchange="peer" (default): the library owns the message buffers and pushes halos / migrants straight
    into the neighbours' buffers over NVLink (CUDA IPC); `connect_dist` distributes the IPC handles once
    (torch.distributed is only the plumbing for that) and a step is a single C call.
    exchange="nccl": the messages live in torch tensors and are exchanged with NCCL send/recv between the
    phases (`DistExchange`)."""

    def __init__(self, rank: int, world: int, rmax: float, *, device: int = 0, capacity: int, halo_cap: int,
                 mig_cap: int, friction=0.85, force=1.0, wrap=True, stream=None, flags: int = 0, exchange: str = "peer", bins: int = 0):
        from . import NativePhysics
        self.rank, self.world, self.wrap, self.exchange_mode = rank, world, wrap, exchange
        self.native = NativePhysics(device=device, precision=N.F32, capacity=capacity, flags=flags, stream=stream, bins=bins)
        self.native.set_settings(rmax, friction, force, wrap)
        self.rmax = rmax
        L = self.native.L
        if exchange == "peer":
            self.native._check(L.plife_slab_configure(self.native.h, rank, world, halo_cap, mig_cap, None))
            return
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        nx = grid_rows(rmax)
        hrec = int(L.plife_slab_halo_records(nx, halo_cap))
        mrec = int(L.plife_slab_migrate_records(mig_cap))
        mk = lambda rec: torch.zeros(rec * 4, dtype=torch.float32, device=self.dev)
        self.halo_send = [mk(hrec), mk(hrec)]
        self.halo_recv = [mk(hrec), mk(hrec)]
        self.mig_send = [mk(mrec), mk(mrec)]
        self.mig_recv = [mk(mrec), mk(mrec)]
        b = N.SlabBuffers()
        for d in range(2):
            b.halo_send[d] = self.halo_send[d].data_ptr()
            b.halo_recv[d] = self.halo_recv[d].data_ptr()
            b.mig_send[d] = self.mig_send[d].data_ptr()
            b.mig_recv[d] = self.mig_recv[d].data_ptr()
        self._bufs = b
        self.native._check(L.plife_slab_configure(self.native.h, rank, world, halo_cap, mig_cap, C.byref(b)))

    # -- peer exchange wiring --
    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.native._check(self.native.L.plife_slab_export(self.native.h, buf))
        return buf.raw

    def connect_ipc(self, down: Optional[bytes], up: Optional[bytes]):
        self._ipc = (C.create_string_buffer(down, 64) if down else None, C.create_string_buffer(up, 64) if up else None)
        self.native._check(self.native.L.plife_slab_connect_ipc(self.native.h, self._ipc[0], self._ipc[1]))

    def connect_dist(self, group=None):
        """All-gather the IPC handles over torch.distributed and map the two ring neighbours.  Collective: if the
        mapping fails on any rank (e.g. CUDA IPC not permitted), every rank raises."""
        import torch
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export_handle(), group=group)
        dn, up = neighbours(self.rank, self.world, True)  # map both ring neighbours; closed boundaries just never use them
        err = None
        try:
            self.connect_ipc(handles[dn] if dn is not None else None, handles[up] if up is not None else None)
        except Exception as e:  # noqa: BLE001 - reported collectively below
            err = e
        ok = torch.tensor([0 if err else 1], device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError(f"peer exchange could not be set up on every rank ({err or 'failed on another rank'})")

    def rows(self):
        lo, hi, nx = C.c_int32(), C.c_int32(), C.c_int32()
        self.native._check(self.native.L.plife_slab_rows(self.native.h, C.byref(lo), C.byref(hi), C.byref(nx)))
        return lo.value, hi.value, nx.value

    def phase(self, which: int, dt: float):
        self.native._check(self.native.L.plife_slab_phase(self.native.h, which, dt))

    def step(self, dt: float, exchange: Optional[DistExchange] = None, nsteps: int = 1):
        if self.exchange_mode == "peer":
            self.native._check(self.native.L.plife_slab_step(self.native.h, dt, nsteps))
            return
        for _ in range(nsteps):
            self.phase(N.SLAB_SORT, dt)
            exchange.exchange(self.halo_send, self.halo_recv, self.wrap, "halo")
            self.phase(N.SLAB_FORCE, dt)
            exchange.exchange(self.mig_send, self.mig_recv, self.wrap, "mig")
            self.phase(N.SLAB_FINISH, dt)

    @property
    def count(self):
        return self.native.count


class VirtualCluster:
    """`world` slabs driven in lockstep inside one process on one device."""

    def __init__(self, world: int, rmax: float, matrix, *, device: int = 0, capacity: int, halo_cap: int, mig_cap: int,
                 wrap=True, friction=0.85, force=1.0, accelerator=(0, ()), exchange: str = "peer", bins: int = 0):
        self.world, self.wrap, self.rmax, self.exchange_mode = world, wrap, rmax, exchange
        self.slabs = [SlabPhysics(r, world, rmax, device=device, capacity=capacity, halo_cap=halo_cap, mig_cap=mig_cap,
                                  wrap=wrap, friction=friction, force=force, exchange=exchange, bins=bins) for r in range(world)]
        for s in self.slabs:
            s.native.set_matrix(matrix)
            s.native.set_accelerator(accelerator[0], accelerator[1])
        self.ex = LocalExchange(world)
        if exchange == "peer":
            for r, s in enumerate(self.slabs):
                dn, up = neighbours(r, world, True)
                s.native._check(s.native.L.plife_slab_connect_local(
                    s.native.h, self.slabs[dn].native.h if dn is not None else None,
                    self.slabs[up].native.h if up is not None else None))

    def upload(self, pos, vel, types, ids=None):
        pos = np.asarray(pos, np.float64).reshape(-1, 2)
        n = pos.shape[0]
        ids = np.arange(n, dtype=np.uint32) if ids is None else np.asarray(ids, np.uint32)
        vel = np.zeros((n, 2)) if vel is None else np.asarray(vel, np.float64)
        own = owner_of_position(pos[:, 1], self.rmax, self.world)
        for r, s in enumerate(self.slabs):
            k = own == r
            s.native.upload(pos[k], vel[k], np.asarray(types)[k], ids[k])

    def step(self, dt: float, nsteps: int = 1):
        if self.exchange_mode == "peer":
            # lockstep over the virtual ranks: every phase is enqueued on all handles before the next one,
            # so a rank spinning on a neighbour's flag always finds the producer already queued
            for _ in range(nsteps):
                for ph in (N.SLAB_SORT, N.SLAB_FORCE, N.SLAB_FINISH):
                    for s in self.slabs:
                        s.phase(ph, dt)
            return
        for _ in range(nsteps):
            for s in self.slabs:
                s.phase(N.SLAB_SORT, dt)
            self._sync()
            self.ex.exchange_all([s.halo_send for s in self.slabs], [s.halo_recv for s in self.slabs], self.wrap)
            self._sync()
            for s in self.slabs:
                s.phase(N.SLAB_FORCE, dt)
            self._sync()
            self.ex.exchange_all([s.mig_send for s in self.slabs], [s.mig_recv for s in self.slabs], self.wrap)
            self._sync()
            for s in self.slabs:
                s.phase(N.SLAB_FINISH, dt)

    def _sync(self):
        # the handles run on their own streams; the copies run on torch's current stream
        import torch
        for s in self.slabs:
            s.native.sync()
        torch.cuda.synchronize()

    def download(self):
        """Concatenation of the slabs in rank order == the single-GPU particle order."""
        parts = [s.native.download() for s in self.slabs]
        from . import Particles
        return Particles(*(np.concatenate([getattr(p, f) for p in parts]) for f in Particles._fields))

    def counts(self):
        return [s.count for s in self.slabs]
