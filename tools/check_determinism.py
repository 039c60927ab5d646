"""Which side is irreproducible?  Runs the 2-GPU slab job and the single-GPU job several times each and
compares the outputs bitwise (by particle id)."""
import os, socket, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import make_state
import test_slab_nccl as T

N, M, RMAX, STEPS = 200_000, 6, 0.01, 15

def slab_run(exchange):
    tmp = tempfile.mkdtemp()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    torch.multiprocessing.spawn(T._worker, args=(2, port, tmp, exchange), nprocs=2, join=True)
    parts = [np.load(os.path.join(tmp, f"rank{r}.npz")) for r in range(2)]
    ids = np.concatenate([p["id"] for p in parts]); pos = np.concatenate([p["pos"] for p in parts])
    o = np.argsort(ids)
    return ids[o], pos[o]

def single_run(b2b):
    import plife
    pos, vel, types, matrix = make_state(N, M, seed=99, vel_scale=0.3, f32=True)
    p = plife.NativePhysics(precision=plife.F32)
    p.set_settings(RMAX, 0.85, 1.0, True); p.set_matrix(matrix); p.upload(pos, vel, types)
    if b2b:
        p.step(0.02, STEPS)
    else:
        for _ in range(STEPS):
            p.step(0.02, 1); p.sync()
    g = p.download(); o = np.argsort(g.id)
    return g.id[o], g.position[o]

def diff(a, b):
    return int((np.abs(a[1] - b[1]).max(axis=1) > 0).sum())

def locate():
    """after few steps: where in each rank's array do the first mismatches sit?"""
    import plife
    os.environ['PLIFE_TEST_STEPS'] = sys.argv[2]
    global STEPS
    STEPS = int(sys.argv[2])
    ref = single_run(True)
    for trial in range(3):
        tmp = tempfile.mkdtemp()
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        torch.multiprocessing.spawn(T._worker, args=(2, port, tmp, "peer"), nprocs=2, join=True)
        for r in range(2):
            p = np.load(os.path.join(tmp, f"rank{r}.npz"))
            ids, pos = p["id"], p["pos"]
            refpos = ref[1][np.searchsorted(ref[0], ids)]
            bad = np.nonzero(np.abs(pos - refpos).max(axis=1) > 0)[0]
            for b in bad[:10]:
                print("   id", ids[b], "idx", b, "slab", pos[b], "ref", refpos[b], "d/ulp", (pos[b] - refpos[b]) / np.spacing(np.float32(refpos[b])), "cell", (pos[b] / RMAX).astype(int), "frac", (pos[b] / RMAX) % 1, flush=True)
            print(f"trial {trial} rank {r}: n={len(ids)} nbad={len(bad)} idx%128 hist:", np.bincount(bad % 128, minlength=128).nonzero()[0][:20], np.bincount(bad % 128, minlength=128).max(), " first idx", bad[:12], flush=True)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "locate":
    locate()
    sys.exit(0)

if __name__ == "__main__":
    s1 = single_run(True); s2 = single_run(True); s3 = single_run(False)
    print("single b2b vs b2b:", diff(s1, s2), " b2b vs synced:", diff(s1, s3), flush=True)
    for ex in ("peer", "nccl"):
        a = slab_run(ex); b = slab_run(ex)
        print(ex, "slab vs slab:", diff(a, b), " slab vs single:", diff(a, s1), diff(b, s1), flush=True)
