"""Independent O(N^2) numpy restatement of one Particle Life step (no cell list).

TEST INFRASTRUCTURE ONLY (see oracle/plife_oracle.c header; parity unpinned).

It follows the *mathematical* definition rather than the reference's loop
structure: every ordered pair (i, j), j != i, with minimum-image distance
0 < d <= rmax contributes.  For nx = floor(1/rmax) >= 3 the reference's 3x3
cell scan visits exactly that set, so the C oracle and this
file must agree up to floating-point reassociation.  Output is in INPUT order
(the reference re-sorts particles by cell; compare via ids).

Formulas: A/Main.java:275-280 (accelerator), B/Range.java:46-57,74-81,89-96
(wrap / wrapConnection / clamp), B/Physics.java:401-402,437,447.
"""
from __future__ import annotations

import numpy as np


def particle_life_force(a, d, beta=0.3):
    """A/Main.java:276-278"""
    return np.where(d < beta, d / beta - 1.0, a * (1.0 - np.abs(1.0 + beta - 2.0 * d) / (1.0 - beta)))


def accel_vectors(kind, params, a, px, py):
    """Builder-defined accelerator catalogue (kind 0 = reference)."""
    d = np.sqrt(px * px + py * py)
    beta = params[0] if len(params) else 0.3
    if kind == 0:
        k = particle_life_force(a, d, beta) / d
        return px * k, py * k
    if kind == 1:
        k = particle_life_force(a, d, beta) / (d * d)
        return px * k, py * k
    if kind == 2:
        k = particle_life_force(a, d, beta) / (d * d * d)
        return px * k, py * k
    if kind == 3:
        k = a * (1.0 - d) / d
        return -py * k, px * k
    if kind == 4:
        ang = -a * np.pi
        c, s = np.cos(ang), np.sin(ang)
        k = (1.0 - d) / d
        return (c * px + s * py) * k, (-s * px + c * py) * k
    if kind == 5:
        r = np.maximum(d, 0.01)
        k = 0.01 / (r * r * r)
        return px * k, py * k
    raise ValueError(kind)


def step(pos, vel, types, matrix, rmax=0.02, friction=0.85, force=1.0, dt=0.02, wrap=True,
         accel_kind=0, accel_params=(0.3,)):
    pos = np.asarray(pos, np.float64).reshape(-1, 2)
    vel = np.asarray(vel, np.float64).reshape(-1, 2)
    types = np.asarray(types, np.int64)
    matrix = np.asarray(matrix, np.float64)
    n = pos.shape[0]
    dx = pos[None, :, 0] - pos[:, None, 0]  # [i, j] = x_j - x_i
    dy = pos[None, :, 1] - pos[:, None, 1]
    if wrap:
        dx = np.where(dx < -0.5, dx + 1.0, np.where(dx >= 0.5, dx - 1.0, dx))
        dy = np.where(dy < -0.5, dy + 1.0, np.where(dy >= 0.5, dy - 1.0, dy))
    d2 = dx * dx + dy * dy
    mask = (d2 != 0) & (d2 <= rmax * rmax)
    np.fill_diagonal(mask, False)
    inv = 1.0 / rmax
    px = np.where(mask, dx * inv, 1.0)
    py = np.where(mask, dy * inv, 0.0)
    a = matrix[types[:, None], types[None, :]]
    ax, ay = accel_vectors(accel_kind, accel_params, a, px, py)
    ax = np.where(mask, ax, 0.0)
    ay = np.where(mask, ay, 0.0)
    k = rmax * force * dt
    mu = friction ** (60.0 * dt)
    nv = np.empty_like(vel)
    nv[:, 0] = vel[:, 0] * mu + ax.sum(axis=1) * k
    nv[:, 1] = vel[:, 1] * mu + ay.sum(axis=1) * k
    npos = nv * dt + pos
    if wrap:
        npos = np.where(npos < 0, npos - np.floor(npos), np.where(npos >= 1, npos - np.floor(npos), npos))
    else:
        npos = np.clip(npos, 0.0, 1.0)
    return npos, nv, mask.sum(axis=1).astype(np.int32)
