"""ctypes loader for the CPU oracle (oracle/plife_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(`plife`) never imports this module.  Parity is UNPINNED by the reference
(it ships no tests and cannot run here); see plife_oracle.c's header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libplife_oracle.so")

ACC_PARTICLE_LIFE = 0
ACC_PARTICLE_LIFE_R = 1
ACC_PARTICLE_LIFE_R2 = 2
ACC_ROTATOR_90 = 3
ACC_ROTATOR_ATTR = 4
ACC_PLANETS = 5


class OracleSettings(C.Structure):
    _fields_ = [
        ("rmax", C.c_double),
        ("friction", C.c_double),
        ("force", C.c_double),
        ("dt", C.c_double),
        ("wrap", C.c_int32),
        ("accel_kind", C.c_int32),
        ("accel_params", C.c_double * 4),
        ("m", C.c_int32),
        ("pad_", C.c_int32),
        ("matrix", C.POINTER(C.c_double)),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "plife_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_diag.argtypes = [vp, C.c_int]
        L.oracle_set_particles.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
        L.oracle_get_particles.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_count.argtypes = [vp]
        L.oracle_count.restype = C.c_int64
        L.oracle_update.argtypes = [vp, C.POINTER(OracleSettings), C.c_int]
        L.oracle_grid.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_get_containers.argtypes = [vp, vp]
        L.oracle_get_neighbor_diag.argtypes = [vp, vp, vp, vp]
        L.oracle_pair_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        for f in ("oracle_range_wrap", "oracle_range_wrap_connection", "oracle_range_clamp"):
            getattr(L, f).argtypes = [C.c_double]
            getattr(L, f).restype = C.c_double
        L.oracle_particle_life_force.argtypes = [C.c_double] * 3
        L.oracle_particle_life_force.restype = C.c_double
        L.oracle_accelerate.argtypes = [C.c_int, vp, C.c_double, vp]
        L.oracle_container_index.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int32)]
        L.oracle_container_index.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Mirror of the reference `Physics` object, CPU fp64 (B/Physics.java)."""

    def __init__(self, rmax=0.02, friction=0.85, force=1.0, dt=0.02, wrap=True, matrix=None,
                 accel_kind=ACC_PARTICLE_LIFE, accel_params=(0.3, 0, 0, 0), threads=1, diag=False):
        self.L = lib()
        self.h = self.L.oracle_create()
        if not self.h:
            raise MemoryError("oracle_create failed")
        self.rmax, self.friction, self.force, self.dt, self.wrap = rmax, friction, force, dt, wrap
        self.accel_kind, self.accel_params = accel_kind, tuple(accel_params)
        self.matrix = None if matrix is None else np.ascontiguousarray(matrix, dtype=np.float64)
        self.threads = threads
        self.L.oracle_set_diag(self.h, 1 if diag else 0)
        self.diag = diag

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n(self):
        return int(self.L.oracle_count(self.h))

    def set_particles(self, pos, vel, types, ids=None):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 2)
        n = pos.shape[0]
        vel = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64).reshape(n, 2)
        types = np.ascontiguousarray(types, dtype=np.int32).reshape(n)
        ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.uint32).reshape(n)
        rc = self.L.oracle_set_particles(self.h, n, _ptr(pos), _ptr(vel), _ptr(types), _ptr(ids))
        if rc:
            raise RuntimeError(f"oracle_set_particles rc={rc}")

    def get_particles(self):
        n = self.n
        pos = np.empty((n, 2), np.float64)
        vel = np.empty((n, 2), np.float64)
        types = np.empty(n, np.int32)
        ids = np.empty(n, np.uint32)
        self.L.oracle_get_particles(self.h, _ptr(pos), _ptr(vel), _ptr(types), _ptr(ids))
        return pos, vel, types, ids

    def update(self, dt=None, threads=None):
        """One `Physics.update()` (B/Physics.java:112)."""
        if dt is not None:
            self.dt = dt
        s = OracleSettings()
        s.rmax, s.friction, s.force, s.dt = self.rmax, self.friction, self.force, self.dt
        s.wrap = 1 if self.wrap else 0
        s.accel_kind = self.accel_kind
        for i in range(4):
            s.accel_params[i] = self.accel_params[i] if i < len(self.accel_params) else 0.0
        s.m = self.matrix.shape[0]
        s.matrix = self.matrix.ctypes.data_as(C.POINTER(C.c_double))
        rc = self.L.oracle_update(self.h, C.byref(s), threads or self.threads)
        if rc:
            raise RuntimeError(f"oracle_update rc={rc}")

    def grid(self):
        nx, ny = C.c_int32(), C.c_int32()
        self.L.oracle_grid(self.h, C.byref(nx), C.byref(ny))
        return nx.value, ny.value

    def containers(self):
        nx, ny = self.grid()
        out = np.empty(nx * ny, np.int32)
        self.L.oracle_get_containers(self.h, _ptr(out))
        return out

    def neighbor_diag(self):
        n = self.n
        cnt = np.empty(n, np.int32)
        hsh = np.empty(n, np.uint64)
        bl = np.empty(n, np.uint8)
        rc = self.L.oracle_get_neighbor_diag(self.h, _ptr(cnt), _ptr(hsh), _ptr(bl))
        if rc:
            raise RuntimeError("neighbor diag not enabled")
        return cnt, hsh, bl

    def pair_stats(self):
        e, h = C.c_int64(), C.c_int64()
        self.L.oracle_pair_stats(self.h, C.byref(e), C.byref(h))
        return e.value, h.value


def accelerate(kind, params, a, pos_xy):
    p = np.array([pos_xy[0], pos_xy[1], 0.0], np.float64)
    prm = np.zeros(4, np.float64)
    prm[: len(params)] = params
    lib().oracle_accelerate(kind, _ptr(prm), float(a), _ptr(p))
    return p[:2].copy()
