"""ctypes binding of libteeflow.so (include/teeflow.h).  Fails loudly when the library is missing: the product
has no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .exceptions import EngineUnavailableError

import os

PKG = Path(__file__).resolve().parent
# TEEFLOW_LIB: alternative build of the same library (kernel tuning experiments); default = the in-tree build
LIB_PATH = Path(os.environ.get("TEEFLOW_LIB", PKG / "libteeflow.so"))

TEEFLOW_MAX_LEVELS = 16
TEEFLOW_U8, TEEFLOW_F32 = 0, 1
ERR_BAD_ARG, ERR_BAD_SHAPE, ERR_CUDA, ERR_NCCL, ERR_STATE = -1, -2, -3, -4, -5


class TeeflowParams(C.Structure):
    _fields_ = [
        ("tau", C.c_double), ("lambda_", C.c_double), ("theta", C.c_double), ("epsilon", C.c_double),
        ("scale_step", C.c_double), ("nscales", C.c_int32), ("warps", C.c_int32),
        ("inner_iterations", C.c_int32), ("outer_iterations", C.c_int32), ("median_filtering", C.c_int32),
        ("max_slots", C.c_int32),
    ]


class TeeflowStats(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int32), ("n_levels", C.c_int32), ("n_slots", C.c_int32), ("grid_ctas", C.c_int32),
        ("solver_launches", C.c_int64), ("kernel_launches", C.c_int64), ("device_ms", C.c_float),
        ("pyramid_ms", C.c_float), ("solver_ms", C.c_float), ("reserved", C.c_float),
        ("double_steps", C.c_int32), ("double_steps_discarded", C.c_int32),
    ]


class TeeflowAnalysis(C.Structure):
    _fields_ = [
        ("mag_hi", C.POINTER(C.c_float)), ("ang_mode", C.POINTER(C.c_float)),
        ("rad_hi", C.POINTER(C.c_double)), ("rad_lo", C.POINTER(C.c_double)),
        ("long_hi", C.POINTER(C.c_double)), ("long_lo", C.POINTER(C.c_double)),
        ("counts", C.POINTER(C.c_int64)),
        ("mag_min", C.c_float), ("mag_max", C.c_float), ("ang_min", C.c_float), ("ang_max", C.c_float),
        ("rad_min", C.c_double), ("rad_max", C.c_double), ("long_min", C.c_double), ("long_max", C.c_double),
    ]


# every symbol include/teeflow.h declares: name -> (restype, argtypes)
_i32p = C.POINTER(C.c_int32)
SIGNATURES = {
    "teeflow_default_params": (None, [C.POINTER(TeeflowParams)]),
    "teeflow_abi_version": (C.c_int, []),
    "teeflow_create": (C.c_int, [C.POINTER(TeeflowParams), C.c_int, C.POINTER(C.c_void_p)]),
    "teeflow_destroy": (C.c_int, [C.c_void_p]),
    "teeflow_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "teeflow_get_param": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]),
    "teeflow_last_error": (C.c_char_p, [C.c_void_p]),
    "teeflow_calc_clip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]),
    "teeflow_calc_clip_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]),
    "teeflow_finish": (C.c_int, [C.c_void_p]),
    "teeflow_calc_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_void_p]),
    "teeflow_calc_clip_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_float, C.c_int]),
    "teeflow_calc_pair_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]),
    "teeflow_get_counters": (C.c_int, [C.c_void_p, _i32p, C.c_int]),
    "teeflow_get_stats": (C.c_int, [C.c_void_p, C.POINTER(TeeflowStats)]),
    "teeflow_get_flow_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "teeflow_level_sizes": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _i32p, _i32p]),
    "teeflow_prepare_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "teeflow_clean_masks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_void_p, C.c_void_p]),
    "teeflow_av_centroids": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p]),
    "teeflow_wase_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "teeflow_set_wase": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "teeflow_get_backgrounds": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "teeflow_analyze_clip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_int,
                                       C.c_int, C.c_double, C.c_double, C.POINTER(TeeflowAnalysis), C.c_void_p]),
    "teeflow_analysis_histogram": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
                                             C.c_void_p]),
    "teeflow_selftest_division": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]),
    "teeflow_saliency_fine_grained": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_void_p]),
    "teeflow_time_launches": (C.c_int, [C.c_void_p, C.c_int]),
    "teeflow_get_launch_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "teeflow_selftest_hypot": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libteeflow.so and type every entry point.  Raises EngineUnavailableError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EngineUnavailableError(
            f"{LIB_PATH} not found: build it with `python -m tee_optical_flow_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise EngineUnavailableError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise EngineUnavailableError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
