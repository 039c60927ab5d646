"""ctypes binding of libplife.so (include/plife.h).  No compute happens here and
there is no fallback: if the CUDA library is missing, importing fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLIFE_LIB") or os.path.join(_HERE, "libplife.so")  # PLIFE_LIB: another build of the same library (A/B timing)

OK = 0
ERR_INVALID, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_STOPPED = -1, -2, -3, -4, -5, -6
F32, F64 = 0, 1
FLAG_UNSTABLE_SORT, FLAG_NO_GRAPH, FLAG_FORCE_V1, FLAG_NO_FUSED_BIN, FLAG_SCAN3, FLAG_NO_CELLS = 1, 2, 4, 8, 32, 64
(ACC_PARTICLE_LIFE, ACC_PARTICLE_LIFE_R, ACC_PARTICLE_LIFE_R2, ACC_ROTATOR_90, ACC_ROTATOR_ATTR,
 ACC_PLANETS) = range(6)
K_BIN, K_SCAN, K_SCATTER, K_GATHER, K_FORCE, K_COUNT = 0, 1, 2, 3, 4, 5
KERNEL_NAMES = ("bin", "scan", "scatter", "gather", "force")


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("capacity", C.c_int64),
                ("flags", C.c_int32), ("bins", C.c_int32), ("stream", C.c_void_p)]


class Settings(C.Structure):
    _fields_ = [("rmax", C.c_double), ("friction", C.c_double), ("force", C.c_double),
                ("wrap", C.c_int32), ("reserved", C.c_int32)]


class StepStats(C.Structure):
    _fields_ = [("n", C.c_int64), ("nx", C.c_int32), ("ny", C.c_int32), ("pair_evals", C.c_int64),
                ("steps", C.c_int64), ("graph_steps", C.c_int64)]


class Cursor(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("size", C.c_double), ("shape", C.c_int32), ("wrap", C.c_int32)]


CURSOR_CIRCLE, CURSOR_SQUARE, CURSOR_INFINITY = 0, 1, 2


class SlabBuffers(C.Structure):
    _fields_ = [("halo_send", C.c_void_p * 2), ("halo_recv", C.c_void_p * 2),
                ("mig_send", C.c_void_p * 2), ("mig_recv", C.c_void_p * 2)]


SLAB_SORT, SLAB_FORCE, SLAB_FINISH = 0, 1, 2


class PlifeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"plife status {status}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python particle-life-app_b200/build.py` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "plife_version": (C.c_int, []),
        "plife_status_string": (C.c_char_p, [C.c_int]),
        "plife_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "plife_destroy": (C.c_int, [vp]),
        "plife_set_settings": (C.c_int, [vp, C.POINTER(Settings)]),
        "plife_get_settings": (C.c_int, [vp, C.POINTER(Settings)]),
        "plife_set_matrix": (C.c_int, [vp, i32, vp]),
        "plife_set_matrix_entry": (C.c_int, [vp, i32, i32, dbl]),
        "plife_get_matrix": (C.c_int, [vp, C.POINTER(i32), vp, i32]),
        "plife_set_accelerator": (C.c_int, [vp, i32, vp, i32]),
        "plife_upload": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "plife_download": (C.c_int, [vp, vp, vp, vp, vp]),
        "plife_download_f32": (C.c_int, [vp, vp, vp, vp]),
        "plife_snapshot_async": (C.c_int, [vp, vp, vp, vp]),
        "plife_snapshot_async_u8": (C.c_int, [vp, vp, vp, vp]),
        "plife_snapshot_wait": (C.c_int, [vp]),
        "plife_init_uniform": (C.c_int, [vp, i64, C.c_uint64]),
        "plife_random_matrix": (C.c_int, [vp, i32, C.c_uint64]),
        "plife_step": (C.c_int, [vp, dbl, i32]),
        "plife_sync": (C.c_int, [vp]),
        "plife_count": (i64, [vp]),
        "plife_type_histogram": (C.c_int, [vp, vp]),
        "plife_request_stop": (C.c_int, [vp]),
        "plife_last_error": (C.c_char_p, [vp]),
        "plife_get_containers": (C.c_int, [vp, vp, i64]),
        "plife_get_step_stats": (C.c_int, [vp, C.POINTER(StepStats)]),
        "plife_debug_neighbors": (C.c_int, [vp, vp, vp]),
        "plife_set_profiling": (C.c_int, [vp, i32]),
        "plife_kernel_times": (C.c_int, [vp, vp, vp]),
        "plife_measure_fp32_peak": (C.c_int, [i32, C.POINTER(dbl)]),
        "plife_device_ptrs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "plife_cursor_count": (C.c_int, [vp, C.POINTER(Cursor), C.POINTER(i64)]),
        "plife_cursor_move": (C.c_int, [vp, C.POINTER(Cursor), dbl, dbl]),
        "plife_cursor_delete": (C.c_int, [vp, C.POINTER(Cursor), C.POINTER(i64)]),
        "plife_append": (C.c_int, [vp, i64, vp, vp, vp]),
        "plife_rebuild": (C.c_int, [vp, i64, vp, vp, vp, i64, vp]),
        "plife_slab_halo_records": (i64, [i32, i64]),
        "plife_slab_migrate_records": (i64, [i64]),
        "plife_slab_configure": (C.c_int, [vp, i32, i32, i64, i64, C.POINTER(SlabBuffers)]),
        "plife_slab_rows": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "plife_slab_phase": (C.c_int, [vp, i32, dbl]),
        "plife_slab_step": (C.c_int, [vp, dbl, i32]),
        "plife_slab_export": (C.c_int, [vp, vp]),
        "plife_slab_connect_ipc": (C.c_int, [vp, vp, vp]),
        "plife_slab_connect_local": (C.c_int, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
