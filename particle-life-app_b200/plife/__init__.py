"""Host-side mirror of the reference's `Physics` object over the C ABI.

`Physics` keeps the reference's surface (B/Physics.java): an accelerator, the
position/type/matrix setter plugins, `settings` (wrap, rmax, friction, force,
dt, matrix) and `update()`; the particle buffers live on the GPU and come back
only through `particles` / `snapshot()` (display-time handoff,
A/PhysicsSnapshot.java:24-62).  All compute is in libplife.so (CUDA, sm_100a);
this module never computes a physics step itself and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import NamedTuple, Optional

import numpy as np

from . import _native as N
from . import synth
from ._native import (ACC_PARTICLE_LIFE, ACC_PARTICLE_LIFE_R, ACC_PARTICLE_LIFE_R2, ACC_PLANETS,
                      ACC_ROTATOR_90, ACC_ROTATOR_ATTR, F32, F64, FLAG_FORCE_V1, FLAG_NO_FUSED_BIN, FLAG_SCAN3, FLAG_NO_CELLS,
                      FLAG_UNSTABLE_SORT, KERNEL_NAMES, PlifeError)

__all__ = ["Physics", "PhysicsSettings", "NativePhysics", "Particles", "PlifeError",
           "DefaultPositionSetter", "DefaultTypeSetter", "DefaultMatrixGenerator", "synth"]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Particles(NamedTuple):
    position: np.ndarray  # [n, 2] float64
    velocity: np.ndarray  # [n, 2] float64
    type: np.ndarray      # [n] int32
    id: np.ndarray        # [n] uint32


class NativePhysics:
    """Thin object wrapper over the plife_* entry points (one handle)."""

    def __init__(self, device: int = 0, precision: int = F32, capacity: int = 0, flags: int = 0,
                 stream: Optional[int] = None, bins: int = 0):
        self.L = N.lib()
        cfg = N.Config(device=device, precision=precision, capacity=capacity, flags=flags, bins=bins,
                       stream=stream)
        h = C.c_void_p()
        rc = self.L.plife_create(C.byref(cfg), C.byref(h))
        if rc != N.OK:
            raise PlifeError(rc, "plife_create: " + self.L.plife_status_string(rc).decode()
                             + " (a CUDA device is required; there is no CPU fallback)")
        self.h = h
        self.precision = precision

    # -- plumbing --
    def _check(self, rc):
        if rc != N.OK:
            raise PlifeError(rc, (self.L.plife_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.plife_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- settings --
    def set_settings(self, rmax, friction, force, wrap):
        s = N.Settings(rmax=rmax, friction=friction, force=force, wrap=1 if wrap else 0, reserved=0)
        self._check(self.L.plife_set_settings(self.h, C.byref(s)))

    def get_settings(self):
        s = N.Settings()
        self._check(self.L.plife_get_settings(self.h, C.byref(s)))
        return dict(rmax=s.rmax, friction=s.friction, force=s.force, wrap=bool(s.wrap))

    def set_matrix(self, matrix):
        m = np.ascontiguousarray(matrix, dtype=np.float64)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError("matrix must be square")
        self._check(self.L.plife_set_matrix(self.h, m.shape[0], _ptr(m)))

    def set_matrix_entry(self, i, j, v):
        self._check(self.L.plife_set_matrix_entry(self.h, i, j, v))

    def get_matrix(self):
        m = C.c_int32()
        self._check(self.L.plife_get_matrix(self.h, C.byref(m), None, 0))
        out = np.empty((m.value, m.value), np.float64)
        self._check(self.L.plife_get_matrix(self.h, C.byref(m), _ptr(out), m.value))
        return out

    def random_matrix(self, m, seed):
        self._check(self.L.plife_random_matrix(self.h, m, seed))

    def set_accelerator(self, kind, params=()):
        p = np.asarray(params, np.float64)
        self._check(self.L.plife_set_accelerator(self.h, kind, _ptr(p) if p.size else None, p.size))

    # -- particles --
    def upload(self, pos, vel, types, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(n, 2)
        types = np.ascontiguousarray(types, dtype=np.int32).reshape(n)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32).reshape(n)
        self._check(self.L.plife_upload(self.h, n, _ptr(pos), _ptr(vel), _ptr(types), _ptr(ids)))

    def download(self) -> Particles:
        n = self.count
        pos = np.empty((n, 2), np.float64)
        vel = np.empty((n, 2), np.float64)
        types = np.empty(n, np.int32)
        ids = np.empty(n, np.uint32)
        self._check(self.L.plife_download(self.h, _ptr(pos), _ptr(vel), _ptr(types), _ptr(ids)))
        return Particles(pos, vel, types, ids)

    def download_f32(self, pos=None, vel=None, types=None):
        """Float snapshot into caller buffers (numpy arrays or raw addresses)."""
        def addr(x):
            if x is None or isinstance(x, int):
                return x
            return x.ctypes.data
        self._check(self.L.plife_download_f32(self.h, addr(pos), addr(vel), addr(types)))

    def snapshot_async(self, pos=None, vel=None, types=None, types_u8=False):
        """Start a float snapshot that overlaps the following steps; `snapshot_wait` completes it.  `types` may be an
        int32 or a uint8 array (addresses: say which with `types_u8`); the one-byte form moves 17 B/particle, not 20."""
        def addr(x):
            if x is None or isinstance(x, int):
                return x
            return x.ctypes.data
        if types is not None and not isinstance(types, int):
            types_u8 = types.dtype == np.uint8
        fn = self.L.plife_snapshot_async_u8 if types_u8 else self.L.plife_snapshot_async
        self._check(fn(self.h, addr(pos), addr(vel), addr(types)))

    def snapshot_wait(self):
        """Completes the OLDEST snapshot requested and not yet waited for (up to two may be in flight)."""
        self._check(self.L.plife_snapshot_wait(self.h))

    def init_uniform(self, n, seed):
        self._check(self.L.plife_init_uniform(self.h, n, seed))

    @property
    def count(self) -> int:
        return int(self.L.plife_count(self.h))

    def type_histogram(self):
        m = self.get_matrix().shape[0]
        out = np.zeros(m, np.int64)
        self._check(self.L.plife_type_histogram(self.h, _ptr(out)))
        return out

    # -- editing (cursor actions of the GUI, on the device) --
    @staticmethod
    def _cursor(x, y, size, shape, wrap):
        return N.Cursor(x=x, y=y, size=size, shape=shape, wrap=1 if wrap else 0)

    def cursor_count(self, x, y, size, shape=N.CURSOR_CIRCLE, wrap=True) -> int:
        out = C.c_int64()
        self._check(self.L.plife_cursor_count(self.h, C.byref(self._cursor(x, y, size, shape, wrap)), C.byref(out)))
        return out.value

    def cursor_move(self, x, y, size, dx, dy, shape=N.CURSOR_CIRCLE, wrap=True):
        self._check(self.L.plife_cursor_move(self.h, C.byref(self._cursor(x, y, size, shape, wrap)), dx, dy))

    def cursor_delete(self, x, y, size, shape=N.CURSOR_CIRCLE, wrap=True) -> int:
        out = C.c_int64()
        self._check(self.L.plife_cursor_delete(self.h, C.byref(self._cursor(x, y, size, shape, wrap)), C.byref(out)))
        return out.value

    def append(self, pos, vel, types):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        k = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(k, 2)
        types = np.ascontiguousarray(types, dtype=np.int32).reshape(k)
        self._check(self.L.plife_append(self.h, k, _ptr(pos), _ptr(vel), _ptr(types)))

    def types(self) -> np.ndarray:
        """Only the types, in array order (4 bytes per particle over PCIe): what the type-count planners need."""
        t = np.empty(self.count, np.int32)
        if t.size:
            self._check(self.L.plife_download_f32(self.h, None, None, t.ctypes.data))
        return t

    def rebuild(self, src, types, place=None, placed=None):
        """Device-side apply of a host-planned rearrangement (plife_rebuild): new particle k = old particle src[k] (or a new
        one for src[k] < 0) with type types[k]; where place[k] >= 0 it is put at placed[place[k]] with zero velocity."""
        src = np.ascontiguousarray(src, dtype=np.int32)
        types = np.ascontiguousarray(types, dtype=np.int32).reshape(src.shape[0])
        place = None if place is None else np.ascontiguousarray(place, dtype=np.int32).reshape(src.shape[0])
        placed = np.zeros((0, 2)) if placed is None else np.ascontiguousarray(placed, dtype=np.float64).reshape(-1, 2)
        self._check(self.L.plife_rebuild(self.h, src.shape[0], _ptr(src), _ptr(types), _ptr(place), placed.shape[0],
                                         _ptr(placed) if placed.shape[0] else None))

    # -- stepping --
    def step(self, dt, nsteps=1):
        self._check(self.L.plife_step(self.h, dt, nsteps))

    def sync(self):
        self._check(self.L.plife_sync(self.h))

    def request_stop(self):
        self.L.plife_request_stop(self.h)

    # -- instrumentation --
    def containers(self):
        st = self.step_stats(pairs=False)
        out = np.empty(st["nx"] * st["ny"], np.int32)
        self._check(self.L.plife_get_containers(self.h, _ptr(out), out.size))
        return out

    def containers_local(self, ncell):
        """Slab mode: END offsets of this rank's LOCAL cells (ghost row below, owned rows, ghost row above), as offsets
        into the rank's sorted array (owned block starts at halo_cap)."""
        out = np.empty(ncell, np.int32)
        self._check(self.L.plife_get_containers(self.h, _ptr(out), out.size))
        return out

    def step_stats(self, pairs=True):
        s = N.StepStats()
        self._check(self.L.plife_get_step_stats(self.h, C.byref(s)))
        return dict(n=s.n, nx=s.nx, ny=s.ny, pair_evals=s.pair_evals, steps=s.steps, graph_steps=s.graph_steps)

    def debug_neighbors(self):
        n = self.count
        cnt = np.empty(n, np.int32)
        hsh = np.empty(n, np.uint64)
        self._check(self.L.plife_debug_neighbors(self.h, _ptr(cnt), _ptr(hsh)))
        return cnt, hsh

    def set_profiling(self, on):
        self._check(self.L.plife_set_profiling(self.h, 1 if on else 0))

    def kernel_times(self):
        ms = np.zeros(N.K_COUNT, np.float64)
        ln = np.zeros(N.K_COUNT, np.int64)
        self._check(self.L.plife_kernel_times(self.h, _ptr(ms), _ptr(ln)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(KERNEL_NAMES)}


# ---------------------------------------------------------------------------
# Setter plugins (host side, vectorised).  Interfaces mirror
# B/PositionSetter.java:5-7, B/TypeSetter.java:5-15, B/MatrixGenerator.java:3-11.
# ---------------------------------------------------------------------------

class DefaultPositionSetter:
    """B/DefaultPositionSetter.java:8-14: x, y ~ U[0,1)."""

    def set(self, types: np.ndarray, n_types: int, rng: np.random.Generator) -> np.ndarray:
        return rng.random((types.shape[0], 2))


class DefaultTypeSetter:
    """B/DefaultTypeSetter.java:8-10: floor(random * nTypes)."""
    uses_state = False  # ignores position / velocity: nothing has to come back from the device for it

    def get_type(self, position, velocity, types, n_types: int, rng: np.random.Generator) -> np.ndarray:
        return np.floor(rng.random(types.shape[0]) * n_types).astype(np.int32)


class DefaultMatrixGenerator:
    """B/DefaultMatrixGenerator.java:6-10 -> DefaultMatrix.randomize(): 2*random - 1."""

    def make_matrix(self, size: int, rng: np.random.Generator) -> np.ndarray:
        return 2.0 * rng.random((size, size)) - 1.0


@dataclass
class PhysicsSettings:
    """B/PhysicsSettings.java:8-38"""
    wrap: bool = True
    rmax: float = 0.02
    friction: float = 0.85
    force: float = 1.0
    dt: float = 0.02
    matrix: np.ndarray = field(default_factory=lambda: np.zeros((6, 6)))

    def deep_copy(self) -> "PhysicsSettings":
        return PhysicsSettings(self.wrap, self.rmax, self.friction, self.force, self.dt, self.matrix.copy())


class Physics:
    """Drop-in for the reference's `Physics` (B/Physics.java) on one B200.

    Differences forced by the device boundary: the accelerator is a kind id
    (a Java/Python lambda cannot run on the GPU), and `particles` is a snapshot
    download rather than a live array; mutate with `set_particles`.
    """

    def __init__(self, accelerator: int = ACC_PARTICLE_LIFE, position_setter=None, matrix_generator=None,
                 type_setter=None, *, particle_count: int = 10000, precision: int = F32, device: int = 0,
                 seed: Optional[int] = None, accelerator_params=(), flags: int = 0, stream=None):
        self.settings = PhysicsSettings()
        self.accelerator = accelerator
        self.accelerator_params = tuple(accelerator_params)
        self.position_setter = position_setter or DefaultPositionSetter()
        self.matrix_generator = matrix_generator or DefaultMatrixGenerator()
        self.type_setter = type_setter or DefaultTypeSetter()
        self.rng = np.random.default_rng(seed)
        self.native = NativePhysics(device=device, precision=precision, flags=flags, stream=stream)
        self._pushed = None
        self.generate_matrix()                     # B/Physics.java:78
        self.set_particle_count(particle_count)    # :79

    # -- step --
    def _push_settings(self):
        s = self.settings
        m = np.asarray(s.matrix, np.float64)
        key = (s.wrap, s.rmax, s.friction, s.force, m.tobytes(), m.shape, self.accelerator, self.accelerator_params)
        if key != self._pushed:
            self.native.set_settings(s.rmax, s.friction, s.force, s.wrap)
            self.native.set_matrix(m)
            self.native.set_accelerator(self.accelerator, self.accelerator_params)
            self._pushed = key

    def update(self):
        """B/Physics.java:112: one simulation step with the current settings."""
        self._push_settings()
        self.native.step(self.settings.dt, 1)

    def force_update_stop(self):  # :149-151
        self.native.request_stop()

    def kill(self):  # :156-158
        self.native.close()

    # -- particles --
    @property
    def particles(self) -> Particles:
        return self.native.download()

    def snapshot(self) -> Particles:
        return self.native.download()

    def set_particles(self, position, velocity, types, ids=None):
        self._push_settings()
        self.native.upload(position, velocity, types, ids)

    @property
    def particle_count(self) -> int:
        return self.native.count

    def _generate(self, n):
        """generateParticle (:290-295): type first, then position; velocity zero (:300-302)."""
        n_types = self.settings.matrix.shape[0]
        zeros = np.zeros((n, 2))
        types = self.type_setter.get_type(zeros, zeros, np.zeros(n, np.int32), n_types, self.rng).astype(np.int32)
        pos = self.ensure_position(self.position_setter.set(types, n_types, self.rng))
        return pos, zeros.copy(), types

    def set_particle_count(self, n: int):
        """B/Physics.java:190-223.  Planned on the host from the types, applied on the device (plife_rebuild): the
        particles never leave the GPU."""
        cur = self.native.count
        if n == cur and cur > 0:
            return
        if cur == 0:
            pos, vel, types = self._generate(n)
            self.set_particles(pos, vel, types)
            return
        self._push_settings()
        old_types = self.native.types()
        if n < cur:  # shuffle first, then keep the first n (:201-210, :278-280)
            keep = self.rng.permutation(cur)[:n]
            self.native.rebuild(keep, old_types[keep])
        else:
            pos, _, types = self._generate(n - cur)
            src = np.concatenate([np.arange(cur), np.full(n - cur, -1)])
            place = np.concatenate([np.full(cur, -1), np.arange(n - cur)])
            self.native.rebuild(src, np.concatenate([old_types, types]), place, pos)

    def set_positions(self):
        """B/Physics.java:166-168: every particle gets a new position (velocity zero); types and ids stay."""
        self._push_settings()
        types = self.native.types()
        pos = self.ensure_position(self.position_setter.set(types, self.settings.matrix.shape[0], self.rng))
        n = len(types)
        self.native.rebuild(np.arange(n), types, np.arange(n), pos)

    def _setter_state(self, idx):
        """Positions / velocities for a TypeSetter that looks at them (B/TypeSetter.java:14); the default one does not."""
        if not getattr(self.type_setter, "uses_state", True):
            z = np.zeros((len(idx), 2))
            return z, z
        p = self.native.download()
        return p.position[idx], p.velocity[idx]

    def set_types(self):
        """B/Physics.java:509-511."""
        self._push_settings()
        types = self.native.types()
        idx = np.arange(len(types))
        pos, vel = self._setter_state(idx)
        t = self.type_setter.get_type(pos, vel, types, self.settings.matrix.shape[0], self.rng)
        self.native.rebuild(idx, np.asarray(t, np.int32))

    def ensure_types(self):
        """B/Physics.java:266-272: particles whose type no longer exists get a new one from the TypeSetter.  Runs against the
        OLD (larger) matrix on the device, so call it before the smaller matrix is pushed (set_matrix_size does)."""
        types = self.native.types()
        m = self.settings.matrix.shape[0]
        bad = np.nonzero(types >= m)[0]
        if len(bad):
            pos, vel = self._setter_state(bad)
            t = types.copy()
            t[bad] = self.type_setter.get_type(pos, vel, types[bad], m, self.rng)
            self.native.rebuild(np.arange(len(types)), t)

    # -- matrix --
    def generate_matrix(self):
        """B/Physics.java:170-176."""
        size = self.settings.matrix.shape[0]
        self.settings.matrix = np.asarray(self.matrix_generator.make_matrix(size, self.rng), np.float64)

    def set_matrix_size(self, new_size: int):
        """B/Physics.java:238-258."""
        prev = self.settings.matrix
        if new_size == prev.shape[0]:
            return
        m = np.asarray(self.matrix_generator.make_matrix(new_size, self.rng), np.float64)
        c = min(prev.shape[0], new_size)
        m[:c, :c] = prev[:c, :c]
        self.settings.matrix = m
        if new_size < prev.shape[0]:
            self.ensure_types()

    def get_type_count(self) -> np.ndarray:
        """A/ExtendedPhysics.java:19-26."""
        self._push_settings()
        return self.native.type_histogram()

    def set_type_count(self, type_count):
        """A/ExtendedPhysics.java:40-118: planned on the host for the whole array at once from the types alone
        (setters.plan_type_count), applied on the device (plife_rebuild)."""
        from .setters import plan_type_count
        m = self.settings.matrix.shape[0]
        want = np.asarray(type_count, np.int64).reshape(-1)
        if want.shape[0] != m:
            raise ValueError(f"Got array of length {want.shape[0]}, but current matrix size is {m}. "
                             "Maybe you should change the matrix size before doing this.")
        self._push_settings()
        src, types, fresh = plan_type_count(self.native.types(), want, self.rng)
        place = np.full(len(src), -1, np.int64)
        place[fresh] = np.arange(int(fresh.sum()))
        # setPosition of every particle that could not be reused or is new (:92-97): new position, zero velocity
        pos = self.ensure_position(self.position_setter.set(types[fresh], m, self.rng)) if fresh.any() else None
        self.native.rebuild(src, types, place, pos)

    def set_type_count_equal(self):
        """A/ExtendedPhysics.java:28-38."""
        from .setters import equal_type_count
        want = equal_type_count(self.native.count, self.settings.matrix.shape[0])
        if want is not None:
            self.set_type_count(want)

    # -- geometry helpers (host, tiny) --
    def ensure_position(self, pos: np.ndarray) -> np.ndarray:
        """B/Physics.java:499-505 with B/Range.java:46-57,89-96."""
        pos = np.asarray(pos, np.float64)
        if self.settings.wrap:
            out = np.where((pos < 0) | (pos >= 1), pos - np.floor(pos), pos)
            return out
        return np.clip(pos, 0.0, 1.0)

    def connection(self, pos1, pos2) -> np.ndarray:
        """B/Physics.java:460-470."""
        d = np.asarray(pos2, np.float64) - np.asarray(pos1, np.float64)
        if self.settings.wrap:
            d = np.where(d < -0.5, d + 1.0, np.where(d >= 0.5, d - 1.0, d))
        return d

    def distance(self, pos1, pos2) -> float:
        """B/Physics.java:477-479."""
        return float(np.sqrt((self.connection(pos1, pos2) ** 2).sum()))
