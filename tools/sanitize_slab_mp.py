import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import plife
from plife.slab import DistExchange, SlabPhysics, owner_of_position
from helpers import make_state
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, m, rmax, steps = 20_000, 6, 0.02, 2
xmode = os.environ.get("PLIFE_XMODE", "peer")
pos, vel, types, matrix = make_state(n, m, seed=99, vel_scale=0.3, f32=True)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sp = SlabPhysics(rank, world, rmax, device=rank, capacity=n, halo_cap=2048, mig_cap=2048, stream=stream.cuda_stream, exchange=xmode)
    if xmode == "peer":
        sp.connect_dist()
    sp.native.set_matrix(matrix)
    own = owner_of_position(pos[:, 1], rmax, world) == rank
    sp.native.upload(pos[own], vel[own], types[own], np.arange(n, dtype=np.uint32)[own])
    sp.step(0.02, DistExchange(rank, world), steps)
    got = sp.native.download()
print("rank", rank, "n", len(got.id), flush=True)
dist.barrier()
dist.destroy_process_group()
