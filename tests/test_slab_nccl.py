"""Two real GPUs, one process each, NCCL send/recv between the slab phases: the distributed run must
reproduce the single-GPU state bit for bit.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import make_state

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, exchange):
    import torch.distributed as dist
    import plife
    from plife.slab import DistExchange, SlabPhysics, owner_of_position
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, m, rmax, steps = 200_000, 6, 0.01, int(os.environ.get('PLIFE_TEST_STEPS', '15'))
        pos, vel, types, matrix = make_state(n, m, seed=99, vel_scale=0.3, f32=True)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            sp = SlabPhysics(rank, world, rmax, device=rank, capacity=n, halo_cap=8192, mig_cap=8192, stream=stream.cuda_stream,
                             exchange=exchange, bins=8)
            if exchange == "peer":
                sp.connect_dist()
            sp.native.set_matrix(matrix)
            own = owner_of_position(pos[:, 1], rmax, world) == rank
            sp.native.upload(pos[own], vel[own], types[own], np.arange(n, dtype=np.uint32)[own])
            sp.step(0.02, DistExchange(rank, world), steps)
            got = sp.native.download()
        lo, hi, nx = sp.rows()
        np.savez(os.path.join(tmp, f"rank{rank}.npz"), pos=got.position, vel=got.velocity, typ=got.type, id=got.id)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_two_gpu_slabs_match_single_gpu(native_lib, tmp_path, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import plife
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    torch.multiprocessing.spawn(_worker, args=(2, port, str(tmp_path), exchange), nprocs=2, join=True)
    n, m, rmax, steps = 200_000, 6, 0.01, 15
    pos, vel, types, matrix = make_state(n, m, seed=99, vel_scale=0.3, f32=True)
    single = plife.NativePhysics(precision=plife.F32, bins=8)  # same internal cell list => same fp32 summation order
    single.set_settings(rmax, 0.85, 1.0, True)
    single.set_matrix(matrix)
    single.upload(pos, vel, types)
    single.step(0.02, steps)
    ref = single.download()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    ids = np.concatenate([p["id"] for p in parts])
    gpos = np.concatenate([p["pos"] for p in parts])
    gvel = np.concatenate([p["vel"] for p in parts])
    assert len(ids) == n
    ir, ig = np.argsort(ref.id), np.argsort(ids)
    assert np.array_equal(ref.id[ir], ids[ig])
    dp = np.abs(ref.position[ir] - gpos[ig]).max(axis=1)
    bad = np.nonzero(dp > 0)[0]
    rows, cnt = np.unique((ref.position[ir][bad, 1] / rmax).astype(int), return_counts=True)
    assert len(bad) == 0, f"{len(bad)} positions differ, max {dp.max():.3g}, rows {dict(zip(rows.tolist(), cnt.tolist()))}"
    assert np.array_equal(ref.velocity[ir], gvel[ig])
