"""The reference app's catalogs of position setters, type setters and matrix generators (SURVEY.md
Appendix B), as vectorised host-side plugins for `plife.Physics`.

The reference calls its plugins once per particle (B/Physics.java:280-330); here one call produces the
whole array, so a 16M-particle re-seed is a handful of numpy expressions and one upload.  Same plugin
protocol as the Default* classes in `plife/__init__.py`:

    position setter   .set(types, n_types, rng)                       -> (n, 2) float64
    type setter       .get_type(position, velocity, types, n_types, rng) -> (n,) int32
    matrix generator  .make_matrix(size, rng)                         -> (size, size) float64

The reference draws from `Math.random()` / `java.util.Random`; streams are not reproducible across
languages, so parity here is distributional (tests check supports, moments and the deterministic parts).
Every position setter works in [-1, 1]^2 and finishes with `p * 0.5 + 0.5`
(A/PositionSetterProvider.java:22-23 and each entry after it); positions may fall outside [0, 1) and
are wrapped or clamped afterwards by `Physics.ensure_position` (B/Physics.java:253-264).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

TAU = 2.0 * np.pi


def _unit_to_world(x, y):
    return np.stack([x * 0.5 + 0.5, y * 0.5 + 0.5], axis=1)


def _polar(angle, radius):
    return _unit_to_world(np.cos(angle) * radius, np.sin(angle) * radius)


class _PositionSetter:
    def __init__(self, name, fn, cite):
        self.name, self._fn, self.__doc__ = name, fn, cite

    def set(self, types, n_types, rng):
        types = np.asarray(types)
        return self._fn(types, int(n_types), rng, len(types))

    def __repr__(self):
        return f"<position setter '{self.name}'>"


def _centered(t, m, rng, n):
    return _unit_to_world(rng.standard_normal(n) * np.float32(0.3), rng.standard_normal(n) * np.float32(0.3))


def _uniform(t, m, rng, n):
    return rng.random((n, 2))


def _uniform_circle(t, m, rng, n):
    return _polar(rng.random(n) * TAU, 0.5 * np.sqrt(rng.random(n)))


def _centered_circle(t, m, rng, n):
    return _polar(rng.random(n) * TAU, 0.5 * rng.random(n))


def _ring(t, m, rng, n):
    return _polar(rng.random(n) * TAU, 0.7 + 0.02 * rng.standard_normal(n))


def _rainbow_ring(t, m, rng, n):
    return _polar((0.3 * rng.standard_normal(n) + t) / m * TAU, 0.7 + 0.02 * rng.standard_normal(n))


def _color_battle(t, m, rng, n):
    centre = t / float(m) * TAU
    angle, radius = rng.random(n) * TAU, rng.random(n) * np.float32(0.1)
    return _unit_to_world(0.5 * np.cos(centre) + np.cos(angle) * radius, 0.5 * np.sin(centre) + np.sin(angle) * radius)


def _color_wheel(t, m, rng, n):
    centre = t / float(m) * TAU
    r0, r1 = float(np.float32(0.3)), float(np.float32(0.2))
    return _unit_to_world(r0 * np.cos(centre) + rng.standard_normal(n) * r1, r0 * np.sin(centre) + rng.standard_normal(n) * r1)


def _line(t, m, rng, n):
    return _unit_to_world(2 * rng.random(n) - 1, (2 * rng.random(n) - 1) * np.float32(0.15))


def _spiral_from(f, rng, n):
    angle = 2 * TAU * f                       # maxRotations = 2
    spread = 0.5 * np.minimum(f, 0.2)
    radius = 0.9 * f + spread * rng.standard_normal(n) * spread
    return _polar(angle, radius)


def _spiral(t, m, rng, n):
    return _spiral_from(rng.random(n), rng, n)


def _rainbow_spiral(t, m, rng, n):
    f = (t + 1) / float(m + 2) + (0.3 / m) * rng.standard_normal(n)
    return _spiral_from(np.clip(f, 0.0, 1.0), rng, n)


_P = "A/PositionSetterProvider.java"
POSITION_SETTERS = OrderedDict((s.name, s) for s in (
    _PositionSetter("centered", _centered, f"{_P}:18-24: gaussian, sigma 0.3f"),
    _PositionSetter("uniform", _uniform, f"{_P}:25 -> B/DefaultPositionSetter.java:8-14"),
    _PositionSetter("uniform circle", _uniform_circle, f"{_P}:26-34: area-uniform disc of radius 0.5"),
    _PositionSetter("centered circle", _centered_circle, f"{_P}:35-43: radius-uniform disc (denser at the centre)"),
    _PositionSetter("ring", _ring, f"{_P}:44-51: radius 0.7 +- 0.02"),
    _PositionSetter("rainbow ring", _rainbow_ring, f"{_P}:52-59: ring with the angle set by the type"),
    _PositionSetter("color battle", _color_battle, f"{_P}:60-71: one small disc per type on a circle of radius 0.5"),
    _PositionSetter("color wheel", _color_wheel, f"{_P}:72-82: one gaussian blob per type on a circle of radius 0.3"),
    _PositionSetter("line", _line, f"{_P}:83-88: horizontal band, half height 0.15"),
    _PositionSetter("spiral", _spiral, f"{_P}:89-100: two turns, spread growing to 0.1 at f = 0.2"),
    _PositionSetter("rainbow spiral", _rainbow_spiral, f"{_P}:101-118: spiral ordered by type"),
))


# -- type setters -------------------------------------------------------------

def map_type(value, n_types):
    """A/TypeSetterProvider.java:39-45: clamp(floor(value * nTypes), 0, nTypes - 1)."""
    v = np.floor(np.asarray(value, np.float64) * n_types)
    return np.clip(v, 0, n_types - 1).astype(np.int32)


class _TypeSetter:
    def __init__(self, name, fn, cite):
        self.name, self._fn, self.__doc__ = name, fn, cite

    def get_type(self, position, velocity, types, n_types, rng):
        types = np.asarray(types, np.int32)
        pos = np.asarray(position, np.float64).reshape(len(types), -1)
        vel = np.asarray(velocity, np.float64).reshape(len(types), -1)
        return np.asarray(self._fn(pos, vel, types, int(n_types), rng), np.int32)

    def __repr__(self):
        return f"<type setter '{self.name}'>"


def _t_random(p, v, t, m, rng):
    return np.floor(rng.random(len(t)) * m)


def _t_randomize10(p, v, t, m, rng):
    pick = rng.random(len(t)) < 0.1
    return np.where(pick, map_type(rng.random(len(t)), m), t)


def _t_slices(p, v, t, m, rng):
    return map_type(p[:, 0], m)


def _t_onion(p, v, t, m, rng):
    c = p[:, :2] - 0.5
    return map_type(np.sqrt((c * c).sum(axis=1)) * 2, m)


def _t_kill_still(p, v, t, m, rng):
    return np.where(np.sqrt((v * v).sum(axis=1)) < 0.01, m - 1, t)


_T = "A/TypeSetterProvider.java"
TYPE_SETTERS = OrderedDict((s.name, s) for s in (
    _TypeSetter("random", _t_random, f"{_T}:14 -> B/DefaultTypeSetter.java:8-10"),
    _TypeSetter("randomize 10%", _t_randomize10, f"{_T}:15-17: each particle re-drawn with probability 0.1"),
    _TypeSetter("slices", _t_slices, f"{_T}:18-20: vertical bands by x"),
    _TypeSetter("onion", _t_onion, f"{_T}:21-23: rings by distance from the centre"),
    _TypeSetter("rotate", lambda p, v, t, m, rng: (t + 1) % m, f"{_T}:24-26"),
    _TypeSetter("flip", lambda p, v, t, m, rng: m - 1 - t, f"{_T}:27-29"),
    _TypeSetter("more of first", lambda p, v, t, m, rng: map_type(rng.random(len(t)) * rng.random(len(t)), m),
                f"{_T}:30-32: product of two uniforms favours low types"),
    _TypeSetter("kill still", _t_kill_still, f"{_T}:33-35: |v| < 0.01 becomes the last type"),
))


# -- matrix generators --------------------------------------------------------

class _MatrixGenerator:
    def __init__(self, name, fn, cite):
        self.name, self._fn, self.__doc__ = name, fn, cite

    def make_matrix(self, size, rng):
        return np.asarray(self._fn(int(size), rng), np.float64).reshape(int(size), int(size))

    def __repr__(self):
        return f"<matrix generator '{self.name}'>"


def _ring_offsets(size):
    """(j - i) mod size for every entry: 0 on the diagonal, 1 / size-1 for the two ring neighbours."""
    i = np.arange(size)
    return (i[None, :] - i[:, None]) % max(size, 1)


def _m_random(size, rng):
    return 2.0 * rng.random((size, size)) - 1.0


def _m_symmetry(size, rng):
    m = _m_random(size, rng)
    low = np.tril(m)                     # upper triangle takes the lower one's values, diagonal kept
    return low + np.tril(m, -1).T


def _neighbours(size):
    d = _ring_offsets(size)
    return (d == 1 % max(size, 1)) | (d == (size - 1) % max(size, 1))


def _m_chains(size, rng):
    d = _ring_offsets(size)
    return np.where((d == 0) | _neighbours(size), 1.0, -1.0)


def _m_chains_n(size, far):
    d = _ring_offsets(size)
    return np.where(d == 0, 1.0, np.where(_neighbours(size), 0.2, far))


def _m_snakes(size, rng):
    m = np.zeros((size, size))
    i = np.arange(size)
    m[i, i] = 1.0
    m[i, (i + 1) % max(size, 1)] = 0.2   # for size 1 this overwrites the diagonal, as the reference does
    return m


_M = "A/MatrixGeneratorProvider.java"
MATRIX_GENERATORS = OrderedDict((g.name, g) for g in (
    _MatrixGenerator("random", _m_random, f"{_M}:16 -> B/DefaultMatrix.java:25-31"),
    _MatrixGenerator("symmetry", _m_symmetry, f"{_M}:17-26: m[i][j] = m[j][i] for j >= i"),
    _MatrixGenerator("chains", _m_chains, f"{_M}:27-39: +1 for self and ring neighbours, -1 otherwise"),
    _MatrixGenerator("chains 2", lambda s, rng: _m_chains_n(s, -1.0), f"{_M}:40-54: 1 / 0.2 / -1"),
    _MatrixGenerator("chains 3", lambda s, rng: _m_chains_n(s, 0.0), f"{_M}:55-69: 1 / 0.2 / 0"),
    _MatrixGenerator("snakes", _m_snakes, f"{_M}:70-77: 1 on the diagonal, 0.2 towards the next type"),
    _MatrixGenerator("zero", lambda s, rng: np.zeros((s, s)), f"{_M}:78"),
))


# -- type-count editing (A/ExtendedPhysics.java) --------------------------------

def _rank_within_type(t):
    """rank[k] = how many earlier entries of `t` have the same value (stable)."""
    if len(t) == 0:
        return np.zeros(0, np.int64)
    by_type = np.argsort(t, kind="stable")
    ts = t[by_type]
    group_start = np.flatnonzero(np.r_[True, ts[1:] != ts[:-1]])
    start_of = np.repeat(group_start, np.diff(np.r_[group_start, len(t)]))
    rank = np.empty(len(t), np.int64)
    rank[by_type] = np.arange(len(t)) - start_of
    return rank


def equal_type_count(n, n_types):
    """A/ExtendedPhysics.java:28-38: ceil(n / nTypes) for every type but the last, which takes the remainder
    (negative when n is small against nTypes: the reference then fails inside setTypeCount, and so does
    `plan_type_count`)."""
    if n_types < 2:
        return None
    c = -(-n // n_types)
    out = np.full(n_types, c, np.int64)
    out[-1] = n - (n_types - 1) * c
    return out


def plan_type_count(types, want, rng):
    """Plan ExtendedPhysics.setTypeCount (A/ExtendedPhysics.java:40-118) for a whole array at once.

    Returns `(src, new_types, fresh)` for the new particle array:
      src[k]        index of the old particle that moves to slot k, or -1 for a brand-new particle
      new_types[k]  its type afterwards
      fresh[k]      True where the reference calls setPosition() again (type changed because the
                    particle could not be reused, or the particle is new)

    The reference shuffles, then sweeps the array once, keeping a particle while its type is still under
    quota and swapping it to the back otherwise; the sweep stops one element short (`while (i < j)`), so
    the last examined particle is never kept.  Which particle the sweep examines next does not depend on
    that particle's type, so on a uniform shuffle this is: walk a random order, keep the first want[t] of
    each type, never the last one.  Left-over slots take the types still short, lowest type first
    (ArrayUtils.findFirstIndexWithLess).  When the total does not change (:101-117) nothing moves and no
    position is re-drawn: the first surplus particles of each over-full type, in shuffled order, are
    re-typed.
    """
    types = np.asarray(types, np.int64)
    want = np.asarray(want, np.int64)
    m, n = len(want), len(types)
    if (want < 0).any():
        raise ValueError("negative type count")
    if n and (types.min() < 0 or types.max() >= m):
        raise ValueError(f"Got array of length {m}, but particles use type {int(types.max())}. "
                         "Maybe you should change the matrix size before doing this.")
    new_n = int(want.sum())
    order = rng.permutation(n)
    t = types[order]
    rank = _rank_within_type(t)
    have = np.bincount(t, minlength=m)
    if new_n == n:
        change = rank < np.maximum(have - want, 0)[t]
        new_types = t.copy()
        new_types[change] = np.repeat(np.arange(m), np.maximum(want - have, 0))
        return order, new_types.astype(np.int32), np.zeros(n, bool)
    keep = rank < want[t]
    if n:
        keep[-1] = False
    kept, rest = order[keep], order[~keep]
    carried = min(new_n, n) - len(kept)              # old particles that survive with a new type
    src = np.concatenate([kept, rest[:carried], np.full(new_n - len(kept) - carried, -1, np.int64)])
    short = want - np.bincount(types[kept], minlength=m)
    new_types = np.concatenate([types[kept], np.repeat(np.arange(m), short)])
    fresh = np.arange(new_n) >= len(kept)
    return src, new_types.astype(np.int32), fresh
