"""Generates tests/golden/*.npz from the CPU oracle.

    python tests/golden/make_golden.py

The reference ships no golden vectors and cannot run here (Java), so these fixtures
are outputs of OUR oracle: they pin the oracle against drift and give the GPU tests committed
input/output pairs.  Inputs are fp32-representable so the same file serves the fp32 and fp64 paths.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "particle-life-app_b200"))

import oracle  # noqa: E402
from plife import synth  # noqa: E402

CASES = {
    "c1_wrap": dict(n=2000, m=6, rmax=0.04, wrap=True, seed=0x5EED0001, steps=3),
    "c1_clamp": dict(n=2000, m=6, rmax=0.04, wrap=False, seed=0x5EED0011, steps=3),
    "fat_cell": dict(n=1500, m=3, rmax=0.065, wrap=True, seed=0x5EED0021, steps=2),
    "nx2_dup": dict(n=400, m=2, rmax=0.5, wrap=True, seed=0x5EED0031, steps=1),
    "rotator": dict(n=1200, m=4, rmax=0.05, wrap=True, seed=0x5EED0041, steps=2, accel_kind=3),
    # half of the particles in a gaussian blob: cells of several hundred particles next to empty ones (what an evolved
    # state looks like); on the GPU the dense CTAs stream their candidates through shared memory in chunks
    "blob": dict(n=8000, m=4, rmax=0.025, wrap=True, seed=0x5EED0051, steps=1),
}


def make(name, c):
    pos, vel, types = synth.uniform_state(c["n"], c["m"], c["seed"])
    u = synth.uniform01(c["seed"] ^ 0x77, np.arange(2 * c["n"], dtype=np.uint64)).reshape(-1, 2)
    vel = (u - 0.5) * 0.08
    pos = pos.astype(np.float32).astype(np.float64)
    vel = vel.astype(np.float32).astype(np.float64)
    if name == "c1_wrap":
        pos[:8, 0] = 1.0  # E1: x == 1.0 is reachable in wrap mode
    if name == "blob":
        g = np.random.default_rng(c["seed"])
        k = c["n"] // 2
        pos[:k] = np.clip(g.normal(0.5, 0.03, (k, 2)), 0, 0.999999).astype(np.float32).astype(np.float64)
    M = synth.random_matrix(c["m"], c["seed"])
    o = oracle.Oracle(rmax=c["rmax"], wrap=c["wrap"], matrix=M, dt=0.02, accel_kind=c.get("accel_kind", 0), diag=True)
    o.set_particles(pos, vel, types)
    out = {}
    for s in range(c["steps"]):
        o.update()
        p, v, t, i = o.get_particles()
        out[f"pos{s + 1}"], out[f"vel{s + 1}"], out[f"typ{s + 1}"], out[f"id{s + 1}"] = p, v, t, i
        if s == 0:
            out["containers1"] = o.containers()
            out["nbr_count1"], out["nbr_hash1"], _ = o.neighbor_diag()
            out["pair_stats1"] = np.array(o.pair_stats(), np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pos0=pos, vel0=vel, typ0=types, matrix=M,
                        rmax=c["rmax"], wrap=c["wrap"], steps=c["steps"], accel_kind=c.get("accel_kind", 0), **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    for k, v in CASES.items():
        if only and k not in only:
            continue
        make(k, v)
        print("wrote", k)
