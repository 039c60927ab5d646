"""Full-size checks (BASELINE configs C2 and C3: 1M and 16M particles) through size-independent properties:
the oracle needs tens of seconds per step at 16M, so parity at these sizes is established by
  * the fp64 GPU path (bit-exact with the oracle at every size the oracle is run on) as the reference for fp32,
  * exact integer invariants: sortedness and stability of the cell-list build, END offsets, the pair-evaluation
    count recomputed from the occupancy histogram, the id multiset,
  * momentum conservation for a symmetric matrix without friction.
"""
import numpy as np
import pytest

import plife
from helpers import rel_l2
from plife import synth

pytestmark = pytest.mark.gpu
DT = 0.02


def make(cfg, precision, seed_shift=0):
    p = plife.NativePhysics(precision=precision)
    p.set_settings(cfg["rmax"], 0.85, 1.0, cfg["wrap"])
    p.random_matrix(cfg["m"], cfg["seed"])
    p.init_uniform(cfg["n"], cfg["seed"] + seed_shift)
    return p


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_full_size_invariants(native_lib, name):
    cfg = synth.CONFIGS[name]
    n, rmax = cfg["n"], cfg["rmax"]
    nx = int(np.floor(1.0 / rmax))
    p = make(cfg, plife.F32)
    p.step(DT, 3)
    # a dt = 0 step rebuilds the cell list and changes nothing else (mu = pow(f, 0) = 1, k2 = 0, x += v*0)
    before = p.download()
    p.step(0.0, 1)
    after = p.download()
    ends = p.containers()
    stats = p.step_stats()
    # id multiset preserved
    assert np.array_equal(np.sort(after.id), np.arange(n, dtype=np.uint32))
    # values untouched by the dt = 0 step, only permuted
    o_b, o_a = np.argsort(before.id), np.argsort(after.id)
    assert np.array_equal(before.position[o_b], after.position[o_a]) and np.array_equal(before.velocity[o_b], after.velocity[o_a])
    # sorted by container, END offsets consistent (B/Physics.java:335-348)
    cx = np.minimum((after.position[:, 0] / rmax).astype(np.int64), nx - 1)
    cy = np.minimum((after.position[:, 1] / rmax).astype(np.int64), nx - 1)
    cell = cx + cy * nx
    assert np.all(np.diff(cell) >= 0)
    occ = np.bincount(cell, minlength=nx * nx)
    assert np.array_equal(ends, np.cumsum(occ)) and ends[-1] == n
    # stable: inside a cell the previous array order is kept
    prev_index = np.empty(n, np.int64)
    prev_index[before.id] = np.arange(n)
    pi = prev_index[after.id]
    same = cell[1:] == cell[:-1]
    assert np.all(pi[1:][same] > pi[:-1][same])
    # pair evaluations = sum_i (occupancy of the 3x3 block around i) - 1, exactly (no particle sits at x == 1.0 here)
    grid = occ.reshape(nx, nx)
    block = sum(np.roll(np.roll(grid, dy, 0), dx, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    assert stats["pair_evals"] == int((grid * block).sum() - n)
    assert (stats["nx"], stats["ny"]) == (nx, nx)


def test_fp32_matches_fp64_gpu_at_16m(native_lib):
    cfg = synth.CONFIGS["C3"]
    a, b = make(cfg, plife.F32), make(cfg, plife.F64)
    # identical fp32-representable start: move the fp32 state into the fp64 handle
    s = a.download()
    b.upload(s.position, s.velocity, s.type, s.id)
    a.step(DT, 1)
    b.step(DT, 1)
    ga, gb = a.download(), b.download()
    assert np.array_equal(ga.id, gb.id)
    assert rel_l2(ga.velocity, gb.velocity) <= 1e-5
    d = np.abs(ga.position - gb.position)
    assert np.minimum(d, 1 - d).max() <= 2e-7


def test_momentum_conservation_symmetric_matrix(native_lib):
    cfg = dict(synth.CONFIGS["C2"])
    M = synth.random_matrix(cfg["m"], 5)
    M = 0.5 * (M + M.T)
    for precision, tol in ((plife.F64, 1e-12), (plife.F32, 2e-4)):
        p = plife.NativePhysics(precision=precision)
        p.set_settings(cfg["rmax"], 1.0, 1.0, True)  # friction 1: no damping
        p.set_matrix(M)
        p.init_uniform(cfg["n"], 77)
        p.step(DT, 5)
        v = p.download().velocity
        # pairwise forces cancel: total momentum stays 0 up to rounding, relative to the sum of |v|
        assert np.abs(v.sum(axis=0)).max() <= tol * np.abs(v).sum()
