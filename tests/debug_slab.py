import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plife
from plife.slab import VirtualCluster
from helpers import make_state
DT=0.02
world, wrap = int(sys.argv[1]) if len(sys.argv)>1 else 2, True
n, m, rmax = 60_000, 5, 0.02
pos, vel, types, matrix = make_state(n, m, seed=31 + world, vel_scale=0.3, f32=True)
single = plife.NativePhysics(precision=plife.F32)
single.set_settings(rmax, 0.85, 1.0, wrap); single.set_matrix(matrix); single.upload(pos, vel, types)
vc = VirtualCluster(world, rmax, matrix, capacity=n, halo_cap=4096, mig_cap=4096, wrap=wrap)
vc.upload(pos, vel, types)
for s in range(6):
    single.step(DT, 1); vc.step(DT, 1)
    ref = single.download(); got = vc.download()
    ir, ig = np.argsort(ref.id), np.argsort(got.id)
    dp = np.abs(ref.position[ir]-got.position[ig]).max(axis=1); dv = np.abs(ref.velocity[ir]-got.velocity[ig]).max(axis=1)
    bad = np.nonzero((dp>0)|(dv>0))[0]
    print("step", s, "counts", vc.counts(), "nbad", len(bad), "max dp", dp.max(), "max dv", dv.max(), "rel dv", dv.max()/np.abs(ref.velocity).max())
    if len(bad):
        y = ref.position[ir][bad,1]; rows = (y/rmax).astype(int)
        print("   rows of bad:", np.unique(rows)[:20], " ... n rows", len(np.unique(rows)))
        print("   sample", bad[:5], dv[bad[:5]], dp[bad[:5]])
