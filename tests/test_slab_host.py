"""Host-side logic of the multi-GPU slab path on CPU: ownership functions and the message routing of
DistExchange under torch.distributed/gloo with world_size 2 and 3 (no GPU, no libplife compute)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plife import slab


def test_ownership_functions_agree():
    for ny in (16, 25, 1000, 2800):
        for world in (1, 2, 3, 4, 8):
            rows = [slab.slab_rows(r, world, ny) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == ny
            for (a, b), (c, d) in zip(rows, rows[1:]):
                assert b == c and b > a
            cy = np.arange(ny)
            own = slab.owner_of_row(cy, world, ny)
            for r, (lo, hi) in enumerate(rows):
                assert np.all(own[lo:hi] == r)


def test_owner_of_position_uses_fp32_rounding_and_clamp():
    rmax, world = 0.02, 4  # ny = 50
    # rows 0-11 -> rank 0, 12-24 -> 1, 25-36 -> 2, 37-49 -> 3; y == 1.0 is clamped into row 49
    y = np.array([0.0, 0.2399, 0.2401, 0.999999, 1.0, 0.9999999999])  # the last rounds to 1.0f
    own = slab.owner_of_position(y, rmax, world)
    assert own.tolist() == [0, 0, 1, 3, 3, 3]


def test_neighbours():
    assert slab.neighbours(0, 1, True) == (None, None)
    assert slab.neighbours(0, 2, True) == (1, 1)
    assert slab.neighbours(0, 2, False) == (None, 1)
    assert slab.neighbours(1, 2, False) == (0, None)
    assert slab.neighbours(0, 4, True) == (3, 1)
    assert slab.neighbours(3, 4, True) == (2, 0)
    assert slab.neighbours(3, 4, False) == (2, None)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, wrap, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = slab.DistExchange(rank, world)
        # payload encodes (sender rank, direction)
        send = [torch.full((8,), 10.0 * rank + 0), torch.full((8,), 10.0 * rank + 1)]
        recv = [torch.full((8,), -1.0), torch.full((8,), -1.0)]
        for _ in range(3):  # repeated exchanges keep their pairing
            ex.exchange(send, recv, wrap)
        out.put((rank, recv[0][0].item(), recv[1][0].item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,wrap", [(2, True), (2, False), (3, True), (3, False)])
def test_dist_exchange_routing_gloo(world, wrap):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, wrap, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        r, below, above = out.get(timeout=120)
        got[r] = (below, above)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        dn, up = slab.neighbours(r, world, wrap)
        # recv[0] = the DOWN neighbour's send[1] (its last row / its up-leavers); recv[1] = the UP neighbour's send[0]
        assert got[r][0] == (10.0 * dn + 1 if dn is not None else -1.0)
        assert got[r][1] == (10.0 * up + 0 if up is not None else -1.0)


# ---------------------------------------------------------------------------
# bench.py's multi-GPU correctness word: pair evaluations implied by the owners' cell histograms (gloo, 2 and 3 ranks)
# ---------------------------------------------------------------------------

def _hist_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from helpers import make_state
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, rmax, first = 6000, 1.0 / 24, 500
        nx = 24
        pos, _, _, _ = make_state(n, 3, seed=17, f32=True)
        pos[:40, 0] = 1.0                      # un-clamped column nx: scans the columns nx-1, 0, 1
        pos[40:50, 1] = 1.0                    # un-clamped row ny on the last slab (slab rule: rows ny-2, ny-1, 0)
        cx = np.minimum((pos[:, 0] / rmax).astype(np.int64), nx - 1)
        cy = np.minimum((pos[:, 1] / rmax).astype(np.int64), nx - 1)
        lo, hi = slab.slab_rows(rank, world, nx)
        mine = (cy >= lo) & (cy < hi)
        # this rank's local cell END offsets: ghost row below | owned rows | ghost row above, owned block starting at `first`
        occ = np.zeros((hi - lo + 2, nx), np.int64)
        np.add.at(occ, (cy[mine] - lo + 1, cx[mine]), 1)
        ends = first + np.cumsum(occ[1:-1].reshape(-1))
        cont = np.concatenate([np.full(nx, first), ends, np.full(nx, ends[-1])]).astype(np.int32)
        got, sane, n_special = bench.histogram_pair_evals(torch, dist, cont, nx, hi - lo + 2, first, rank, world, True,
                                                          pos[mine].astype(np.float32), rmax, lo, device="cpu")
        out.put((rank, got, sane, n_special, int(mine.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_histogram_pair_evals_matches_the_oracle_count(world):
    """Sum over the ranks == the oracle's candidate-pair count of the whole state, including particles sitting exactly on
    x == 1.0; the particles on y == 1.0 follow the slab rule (plife_internal.h: scan_row), accounted for separately."""
    import oracle
    from helpers import make_state
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hist_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[2] for r in res) and sum(r[4] for r in res) == 6000 and sum(r[3] for r in res) >= 50
    n, rmax, nx = 6000, 1.0 / 24, 24
    pos, vel, types, matrix = make_state(n, 3, seed=17, f32=True)
    pos[:40, 0] = 1.0
    pos[40:50, 1] = 1.0
    o = oracle.Oracle(rmax=rmax, matrix=matrix, wrap=True, dt=0.0)
    o.set_particles(pos, vel, types)
    o.update()
    total = o.pair_stats()[0]
    # the reference scans the rows ny-1, 0, 1 for a particle on y == 1.0, a slab the rows ny-2, ny-1, 0: difference of the two
    cx = np.minimum((pos[:, 0] / rmax).astype(np.int64), nx - 1)
    cy = np.minimum((pos[:, 1] / rmax).astype(np.int64), nx - 1)
    occ = np.zeros((nx, nx), np.int64)
    np.add.at(occ, (cy, cx), 1)
    adjust = 0
    for i in np.nonzero(pos[:, 1] == 1.0)[0]:
        c0 = int(pos[i, 0] / rmax)
        cols = [(c0 + d) % nx if c0 < nx else [nx - 1, 0, 1][d + 1] for d in (-1, 0, 1)]
        adjust += int(occ[nx - 2, cols].sum() - occ[1, cols].sum())
    assert sum(r[1] for r in res) == total + adjust
