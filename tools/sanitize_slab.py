import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tools/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plife
from plife.slab import VirtualCluster
from helpers import make_state
n, m, rmax = 20_000, 5, 0.02
pos, vel, types, matrix = make_state(n, m, seed=33, vel_scale=0.3, f32=True)
vc = VirtualCluster(2, rmax, matrix, capacity=n, halo_cap=2048, mig_cap=2048, wrap=True)
vc.upload(pos, vel, types)
vc.step(0.02, 3)
print("counts", vc.counts())
single = plife.NativePhysics(precision=plife.F32)
single.set_settings(rmax, 0.85, 1.0, True); single.set_matrix(matrix); single.upload(pos, vel, types); single.step(0.02, 3)
print(single.step_stats())
