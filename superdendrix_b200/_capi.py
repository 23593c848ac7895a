"""ctypes binding of libsdx.so — one prototype per entry point of include/sdx.h.

There is no CPU fallback: if the library is missing, or a function fails, the
caller gets an exception (``SdxError``), never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsdx.so")

SDX_OK = 0
SDX_ERR_INVALID = -1
SDX_ERR_CUDA = -2
SDX_ERR_STATE = -3
SDX_ERR_UNSUPPORTED = -4
SDX_MAX_K = 8
SDX_STAT_FIXED = 0
SDX_STAT_MAXK = 1


class SdxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__("libsdx error %d: %s" % (code, message))
        self.code = code
        self.message = message


class Hit(C.Structure):
    _fields_ = [("W", C.c_double), ("features", C.c_int32 * 3), ("coverage", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("plan_ms", C.c_float), ("trade_ms", C.c_float), ("score_ms", C.c_float),
                ("total_ms", C.c_float), ("launches", C.c_int32), ("trade_launches", C.c_int32)]


_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_ctx = C.c_void_p

# name -> (restype, argtypes); the list IS the export contract checked by tests/test_capi_symbols.py
PROTOTYPES = {
    "sdx_version": (C.c_int, []),
    "sdx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sdx_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "sdx_destroy": (C.c_int, [_ctx]),
    "sdx_last_error": (C.c_char_p, [_ctx]),
    "sdx_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "sdx_get_timing": (C.c_int, [_ctx, C.POINTER(Timing)]),
    "sdx_total_launches": (C.c_int, [_ctx, _i64p]),
    "sdx_upload_matrix": (C.c_int, [_ctx, _u32p, C.c_int, C.c_int, C.c_int]),
    "sdx_upload_weights": (C.c_int, [_ctx, _f64p, _u32p, C.c_int, C.c_int]),
    "sdx_score_sets": (C.c_int, [_ctx, _i32p, _i32p, C.c_int64, C.c_int, _f64p, _i32p, _i32p, _i32p]),
    "sdx_union_sums": (C.c_int, [_ctx, _i32p, _i32p, C.c_int64, C.c_int, _f64p]),
    "sdx_feature_scores": (C.c_int, [_ctx, C.c_int, _f64p, _f64p, _i32p, _i32p]),
    "sdx_trade_shape": (C.c_int, [_ctx, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdx_curveball": (C.c_int, [_ctx, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, _u32p, _i64p]),
    "sdx_curveball_features": (C.c_int, [_ctx, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, _u32p]),
    "sdx_set_burn_in": (C.c_int, [_ctx, C.c_int64, C.c_int]),
    "sdx_set_pair_group": (C.c_int, [_ctx, C.c_int]),
    "sdx_pool_begin": (C.c_int, [_ctx, C.c_uint64, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "sdx_pool_build": (C.c_int, [_ctx, C.c_int, C.c_int]),
    "sdx_pool_commit": (C.c_int, [_ctx]),
    "sdx_pool_valid": (C.c_int, [_ctx, C.c_uint64, C.c_int64]),
    "sdx_curveball_replay": (C.c_int, [_ctx, _u32p, _i32p, _i32p, _u32p, C.c_int64, _u32p, _i64p]),
    "sdx_null_pvalue": (C.c_int, [_ctx, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _i32p,
                                  C.c_int, _f64p, _i64p, _f64p, _f64p]),
    "sdx_null_pvalue_rows": (C.c_int, [_ctx, _u32p, C.c_int64, C.c_int, _i32p, C.c_int, _f64p, _i64p, _f64p, _f64p]),
    "sdx_scan": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.POINTER(Hit), _i64p]),
    "sdx_scan_range": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Hit), _i64p]),
    "sdx_condperm": (C.c_int, [_ctx, C.c_int, _i32p, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, _i64p, _f64p]),
    "sdx_condperm_batch": (C.c_int, [_ctx, _i32p, _i32p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, _i64p, _f64p]),
    "sdx_ranksum": (C.c_int, [_ctx, _i32p, _i32p, C.c_int64, C.c_int, _f64p, _f64p, _i64p, _i32p]),
    "sdx_write_nulls_tsv": (C.c_int, [_u32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                      C.c_char_p, C.c_int64, C.c_int]),
    "sdx_read_nulls_tsv": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p),
                                     C.POINTER(C.c_char_p), _u32p, _i64p, C.c_int]),
    "sdx_io_last_error": (C.c_char_p, []),
    "sdx_measure_int_peaks": (C.c_int, [_ctx, _f64p, _f64p]),
    "sdx_test_philox": (C.c_int, [_ctx, _u32p, _u32p, C.c_int, _u32p]),
    "sdx_test_trade_pairs": (C.c_int, [_ctx, C.c_uint64, C.c_uint64, C.c_int, C.c_int, _i32p, _i32p]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libsdx.so and attach the prototypes.  Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdxError(SDX_ERR_STATE, "%s is not built (run `python -m superdendrix_b200.build`); "
                           "there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)       # AttributeError = missing export: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
