"""ctypes binding of libb200kf.so (include/b200kf.h).

The shared library is built in-tree by csrc/build.py (nvcc, sm_100a).  There is no CPU
implementation behind this module: if the library is missing `load()` raises, and if no CUDA
device is visible every compute call raises `B200KFError` (status B200KF_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t,
                    c_uint8, c_uint32, c_uint64, c_void_p)

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200KF_LIB") or os.path.join(PKG_DIR, "libb200kf.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4
MODEL_CV6, MODEL_CA8, MODEL_BIAS9, MODEL_CA15, MODEL_CVBU6 = 0, 1, 2, 3, 4
F64, F32 = 0, 1
EV_GPS, EV_IMU, EV_PREDICT_ONLY = 0, 1, 2
RUN_AUTO, RUN_FORCE_DENSE, RUN_FORCE_DECOUPLED, RUN_AGG_CONTINUE, RUN_FORCE_SYMMETRIC = 0, 1, 2, 4, 8
RUN_FORCE_SHARED = 16
RUN_SMALL_TILES = 32


class ModelDesc(Structure):
    _fields_ = [("model_id", c_int32), ("precision", c_int32), ("q", c_double * 16),
                ("r_gps", c_double * 4), ("r_imu", c_double * 16), ("accel_coef", c_double)]


class Best(Structure):
    _fields_ = [("argmin_sum_score", c_double), ("argmin_sum_index", c_uint64),
                ("first_feasible_key", c_uint64), ("first_feasible_index", c_uint64),
                ("n_feasible", c_uint64), ("n_near_threshold", c_uint64)]


class BestProduct(Structure):
    _fields_ = [("score", c_double), ("index", c_uint64), ("runner_up_score", c_double), ("n_scored", c_uint64)]


class B200KFError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libb200kf status {status}: {message}")
        self.status = status


# name -> (restype, argtypes); every symbol declared in include/b200kf.h
SIGNATURES = {
    "b200kf_version": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "b200kf_last_error": (c_size_t, [c_char_p, c_size_t]),
    "b200kf_device_count": (c_int, [POINTER(c_int)]),
    "b200kf_model_default_desc": (c_int, [c_int32, c_int32, POINTER(ModelDesc)]),
    "b200kf_model_create": (c_int, [POINTER(ModelDesc), POINTER(c_void_p)]),
    "b200kf_model_destroy": (c_int, [c_void_p]),
    "b200kf_model_dims": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "b200kf_run_batch": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_double, c_uint32, c_void_p]),
    "b200kf_run_batch_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_double, c_uint32, c_void_p]),
    "b200kf_run_batch_gated": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_uint32, c_void_p]),
    "b200kf_run_batch_host": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_uint32]),
    "b200kf_score_schedules": (c_int, [c_void_p, c_int32, POINTER(c_uint8), POINTER(c_double), c_double,
                                       c_double, POINTER(c_double), c_uint64, c_uint64, c_double,
                                       c_double, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_score_schedules_strided": (c_int, [c_void_p, c_int32, POINTER(c_uint8), POINTER(c_double), c_double,
                                               c_double, POINTER(c_double), c_uint64, c_int32, c_uint64, c_uint64,
                                               c_double, c_double, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_run_masked": (c_int, [c_void_p, c_int32, POINTER(c_uint8), POINTER(c_double), POINTER(c_double),
                                  c_double, c_double, POINTER(c_double), POINTER(c_double), c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_score_products": (c_int, [c_void_p, c_int32, POINTER(c_int32), c_void_p, c_void_p, c_void_p, c_void_p,
                                      POINTER(c_double), POINTER(c_double), c_uint64, c_uint64, c_void_p,
                                      c_void_p, c_uint32, c_void_p]),
    "b200kf_track_lerp": (c_int, [c_int64, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "b200kf_rmse_vs_track": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "b200kf_select_greedy": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_int64, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_select_greedy_gains": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                           c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_select_greedy_chunk": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_uint32,
                                           c_void_p]),
    "b200kf_gain_record_bytes": (c_int, [c_void_p, POINTER(c_size_t)]),
    "b200kf_run_gains": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int32, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b200kf_cov_update": (c_int, [c_int32, c_int32, c_int32, POINTER(c_int32), POINTER(c_double),
                                  POINTER(c_double), c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200kf_fill_payload": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_uint64, c_void_p, c_void_p]),
    "b200kf_probe_fma_peak": (c_int, [c_int32, POINTER(c_double)]),
}

_lib = None


def load():
    """Load libb200kf.so and declare the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python sensorfusion-kalmanfilter_b200/csrc/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    lib = load()
    buf = ctypes.create_string_buffer(512)
    lib.b200kf_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(status: int):
    if status != OK:
        raise B200KFError(status, last_error())


def device_count() -> int:
    lib = load()
    c = c_int(0)
    check(lib.b200kf_device_count(byref(c)))
    return c.value


def version():
    lib = load()
    a, b = c_int(0), c_int(0)
    check(lib.b200kf_version(byref(a), byref(b)))
    return a.value, b.value
