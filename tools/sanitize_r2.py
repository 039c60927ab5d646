"""Driver for compute-sanitizer over the code paths added in round 2 (fine bins, four-wide staging with masked and unmasked walks, reference-slot rank,
one-launch scan, small_sort, warp-per-cell kernel, graph replay, plife_rebuild, host-free slab step with the side stream):
compute-sanitizer --tool memcheck python tools/sanitize_r2.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import plife
from plife.slab import VirtualCluster
from helpers import make_state

rng = np.random.default_rng(7)


def run(n, m, rmax, wrap, flags=0, bins=0, steps=3, accel=(0, (0.3,)), precision=plife.F32, borders=True, band=False):
    pos, vel, types, matrix = make_state(n, m, seed=n + m, vel_scale=0.2, f32=True)
    if band:  # grid rows with empty ends: the CTAs straddle rows, the staged kernel walks with masks
        pos[:, 0] = 0.41 + 0.05 * pos[:, 0]
        pos = pos.astype(np.float32).astype(np.float64)
    if borders:
        pos[:10, 0] = 1.0
        pos[10:20, 1] = 1.0
    p = plife.NativePhysics(precision=precision, flags=flags, bins=bins)
    p.set_settings(rmax, 0.85, 1.0, wrap)
    p.set_matrix(matrix)
    p.set_accelerator(*accel)
    p.upload(pos, vel, types)
    p.step(0.02, steps)
    out = p.download()
    assert np.isfinite(out.position).all() and np.isfinite(out.velocity).all()
    assert p.step_stats()["pair_evals"] > 0
    p.containers()
    p.close()


for wrap in (True, False):
    run(9000, 6, 0.04, wrap, steps=6)                              # small_sort + warp-per-cell kernel + graph replay
    run(9000, 6, 0.04, wrap, flags=plife.FLAG_NO_CELLS, steps=6)   # small_sort + staged kernel (fine bins, bulk copies)
    run(9000, 3, 0.04, wrap, accel=(3, ()))                        # literal visitor in the warp-per-cell kernel
    run(70000, 5, 0.012, wrap, bins=8)                             # one-launch scan (several tiles), staged kernel, reference slots
    run(70000, 5, 0.012, wrap, bins=2, flags=plife.FLAG_SCAN3)
    run(70000, 5, 0.04, wrap)                                      # 112 particles per cell: chunked staging with fine bins
    run(30000, 5, 0.02, wrap, flags=plife.FLAG_NO_CELLS, bins=8, band=True, borders=False)  # masked walk (rows with empty ends)
    run(30000, 4, 0.02, wrap, flags=plife.FLAG_NO_CELLS, bins=4, accel=(3, ()))              # literal visitor in the staged kernel
    run(70000, 5, 0.004, wrap)                                     # one particle per cell: v1 kernel
    run(20000, 4, 0.03, wrap, precision=plife.F64)
# plife_rebuild
p = plife.Physics(particle_count=12000, seed=3)
p.update(); p.set_type_count_equal(); p.set_type_count([10, 20, 30, 40, 50, 9000]); p.set_particle_count(4000); p.set_particle_count(5000)
p.set_matrix_size(3); p.update()
assert p.particle_count == 5000
# slab step: side stream, migration exchange next to the interior launch, device-resident counts
n, m, rmax = 30_000, 5, 0.02
pos, vel, types, matrix = make_state(n, m, seed=33, vel_scale=0.3, f32=True)
pos[:20, 1] = 1.0
for world in (2, 3):
    vc = VirtualCluster(world, rmax, matrix, capacity=n, halo_cap=4096, mig_cap=2048, wrap=True, bins=4)
    vc.upload(pos, vel, types)
    vc.step(0.02, 5)
    assert sum(vc.counts()) == n
    vc.slabs[0].native.step_stats()
print("round-2 paths ok")
