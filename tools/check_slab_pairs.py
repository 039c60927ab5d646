"""Debug probe: per-slab pair-evaluation counts against the cell histograms of the owning slabs (virtual ranks, one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plife
from plife.slab import VirtualCluster
from helpers import make_state

world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n, m, rmax, steps = 200_000, 6, 0.01, 6
pos, vel, types, matrix = make_state(n, m, seed=31, vel_scale=0.3, f32=True)
halo_cap = 8192
vc = VirtualCluster(world, rmax, matrix, capacity=n, halo_cap=halo_cap, mig_cap=8192, wrap=True, bins=8)
vc.upload(pos, vel, types)
single = plife.NativePhysics(bins=8)
single.set_settings(rmax, 0.85, 1.0, True); single.set_matrix(matrix); single.upload(pos, vel, types)
vc.step(0.02, steps); single.step(0.02, steps)
tot = 0
occs = []
for s in vc.slabs:
    lo, hi, nx = s.rows()
    nly = hi - lo + 2
    cont = s.native.containers_local(nx * nly).astype(np.int64).reshape(nly, nx)
    ends = cont[1:nly - 1]
    occ = np.diff(np.concatenate([[halo_cap], ends.reshape(-1)])).reshape(nly - 2, nx)
    ghost_lo = np.diff(np.concatenate([[cont[0, 0] - 0], cont[0]]))  # unknown start: only differences inside the row
    occs.append((occ, cont))
for r, s in enumerate(vc.slabs):
    occ, cont = occs[r]
    below = occs[(r - 1) % world][0][-1]
    above = occs[(r + 1) % world][0][0]
    full = np.concatenate([below[None], occ, above[None]])
    row3 = full + np.roll(full, 1, axis=1) + np.roll(full, -1, axis=1)
    nine = row3[:-2] + row3[1:-1] + row3[2:]
    expect = int((occ * nine).sum() - occ.sum())
    st = s.native.step_stats()
    # ghost rows as this slab holds them
    gl = np.diff(cont[0]); ga = np.diff(cont[-1])
    print(r, "rows", s.rows()[:2], "n", st["n"], "pair_evals", st["pair_evals"], "expected", expect, "diff", st["pair_evals"] - expect,
          "| ghost-below matches owner:", np.array_equal(gl, below[1:]), "ghost-above:", np.array_equal(ga, above[1:]),
          "ghost sizes", cont[0, -1] - (cont[0, 0] - below[0]), int(below.sum()), cont[-1, -1] - cont[-2, -1], int(above.sum()))
    tot += st["pair_evals"]
print("sum", tot, "single", single.step_stats()["pair_evals"])
