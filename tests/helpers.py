"""Shared test helpers: seeded states, oracle runs, comparison norms."""
import numpy as np

import oracle
from plife import synth


def make_state(n, m, seed, vel_scale=0.0, f32=True):
    pos, vel, types = synth.uniform_state(n, m, seed)
    if vel_scale:
        u = synth.uniform01(seed ^ 0xABCDEF, np.arange(2 * n, dtype=np.uint64)).reshape(n, 2)
        vel = (u - 0.5) * 2.0 * vel_scale
    matrix = synth.random_matrix(m, seed)
    if f32:  # the fp32 path stores fp32: give BOTH sides the rounded values
        pos = pos.astype(np.float32).astype(np.float64)
        vel = vel.astype(np.float32).astype(np.float64)
    return pos, vel, types, matrix


def oracle_step(pos, vel, types, matrix, ids=None, steps=1, diag=False, threads=1, **kw):
    o = oracle.Oracle(matrix=matrix, diag=diag, threads=threads, **kw)
    o.set_particles(pos, vel, types, ids)
    for _ in range(steps):
        o.update()
    return o


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_over_rms(a, b):
    return float(np.abs(a - b).max() / max(np.sqrt(np.mean(b * b)), 1e-300))
