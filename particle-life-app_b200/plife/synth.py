"""Seeded synthetic initial states (host side).

The reference's default setters are unseeded `Math.random()` calls
(B/DefaultPositionSetter.java:8-14, B/DefaultTypeSetter.java:8-10,
B/DefaultMatrix.java:25-31), so they are not reproducible.  These generators
have the same distributions (x, y ~ U[0,1); type = floor(u * nTypes);
matrix = 2u - 1) but draw from a counter-based SplitMix64 stream so that the
host (numpy), the device generator (`plife_init_uniform`, csrc/plife_api.cu)
and the test oracle all see bit-identical states.

Stream layout for particle i under `seed`: draw(4i) -> x, draw(4i+1) -> y,
draw(4i+2) -> type; matrix entry k under `seed ^ MATRIX_SALT`: draw(k).
draw(c) = mix(seed + (c+1) * GOLDEN); u = (draw >> 11) * 2^-53.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
MATRIX_SALT = 0x4D41545249583634  # "MATRIX64"
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def draw(seed: int, counters: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        c = counters.astype(np.uint64) + np.uint64(1)
        return _mix(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + c * GOLDEN)


def uniform01(seed: int, counters: np.ndarray) -> np.ndarray:
    return (draw(seed, counters) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_state(n: int, n_types: int, seed: int, chunk: int = 1 << 22):
    """positions U[0,1)^2 (fp64), zero velocities, types U{0..n_types-1}."""
    pos = np.empty((n, 2), np.float64)
    types = np.empty(n, np.int32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        i = np.arange(s, e, dtype=np.uint64) * np.uint64(4)
        pos[s:e, 0] = uniform01(seed, i)
        pos[s:e, 1] = uniform01(seed, i + np.uint64(1))
        types[s:e] = np.floor(uniform01(seed, i + np.uint64(2)) * n_types).astype(np.int32)
    vel = np.zeros((n, 2), np.float64)
    return pos, vel, types


def random_matrix(m: int, seed: int) -> np.ndarray:
    """U[-1,1) row-major m x m (B/DefaultMatrix.java:25-31)."""
    k = np.arange(m * m, dtype=np.uint64)
    return (2.0 * uniform01(seed ^ MATRIX_SALT, k) - 1.0).reshape(m, m)


# The concrete configs of SURVEY.md 8(d).  seed = 0x5EED0000 + config index.
CONFIGS = {
    "C1": dict(n=10_000, m=6, rmax=0.04, wrap=True, seed=0x5EED0001),
    "C1d": dict(n=10_000, m=6, rmax=0.02, wrap=True, seed=0x5EED0001),
    "C2": dict(n=1_000_000, m=8, rmax=0.004, wrap=True, seed=0x5EED0002),
    "C3": dict(n=16_000_000, m=16, rmax=0.001, wrap=True, seed=0x5EED0003),
    "C3lo": dict(n=16_000_000, m=16, rmax=2.0 ** -12, wrap=True, seed=0x5EED0003),
    "C4": dict(n=128_000_000, m=8, rmax=1.0 / 2800.0, wrap=True, seed=0x5EED0004),
    "C5": dict(n=4_000_000, m=8, rmax=0.002, wrap=False, seed=0x5EED0005),
}
