"""Loader of the CUDA shared library (C ABI: include/similarity_transform.h).

There is no CPU fallback: if libsimilarity_transform.so is missing and cannot be built,
importing the solver fails loudly.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

c_f32p = ctypes.POINTER(ctypes.c_float)
c_u32p = ctypes.POINTER(ctypes.c_uint32)


class StOptions(ctypes.Structure):
    """st_options (include/similarity_transform.h)."""
    _fields_ = [
        ("eps", ctypes.c_float),
        ("max_iter", ctypes.c_uint32),
        ("form", ctypes.c_int32),
        ("sweep", ctypes.c_int32),
        ("threads", ctypes.c_int32),
        ("ctas", ctypes.c_int32),
        ("kernel", ctypes.c_int32),
        ("l2_keep_pct", ctypes.c_int32),
        ("stop", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
    ]


class StResult(ctypes.Structure):
    """st_result (include/similarity_transform.h)."""
    _fields_ = [
        ("eigen_val", ctypes.c_float),
        ("iter_count", ctypes.c_uint32),
        ("passes", ctypes.c_uint32),
        ("launches", ctypes.c_uint32),
        ("loop_ms", ctypes.c_float),
        ("total_ms", ctypes.c_float),
        ("round_us_median", ctypes.c_float),
        ("round_us_min", ctypes.c_float),
        ("bytes_per_round", ctypes.c_uint64),
        ("status", ctypes.c_int32),
        ("grid", ctypes.c_uint32),
        ("kernel_id", ctypes.c_uint32),
        ("threads", ctypes.c_uint32),
    ]


class StStreamPlan(ctypes.Structure):
    """st_stream_plan (include/similarity_transform.h)."""
    _fields_ = [
        ("block_rows", ctypes.c_uint32),
        ("blocks", ctypes.c_uint32),
        ("slots", ctypes.c_uint32),
        ("streamed", ctypes.c_uint32),
        ("cache_bytes", ctypes.c_uint64),
        ("h2d_bytes_first", ctypes.c_uint64),
        ("h2d_bytes_per_round", ctypes.c_uint64),
        ("h2d_bytes_total", ctypes.c_uint64),
    ]


# every symbol include/similarity_transform.h declares: name -> (restype, argtypes)
_VP = ctypes.c_void_p
SYMBOLS = {
    # Part 1 -- reference wrapper/similarity_transform.cpp:3-37
    "make_queue": (None, [ctypes.POINTER(_VP)]),
    "max_eigen_value": (ctypes.c_int64, [_VP, _VP, _VP, _VP, ctypes.c_uint, _VP]),
    # Part 2
    "st_last_error": (ctypes.c_char_p, []),
    "st_device_count": (ctypes.c_int, []),
    "st_default_options": (None, [ctypes.POINTER(StOptions)]),
    "st_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_VP)]),
    "st_destroy": (None, [_VP]),
    "st_device_info": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t),
                                      ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]),
    "st_malloc": (ctypes.c_int, [_VP, ctypes.c_size_t, ctypes.POINTER(_VP)]),
    "st_free": (ctypes.c_int, [_VP, _VP]),
    "st_memcpy_h2d": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "st_memcpy_d2h": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "st_synchronize": (ctypes.c_int, [_VP]),
    "st_pin_host": (ctypes.c_int, [_VP, _VP, ctypes.c_size_t]),
    "st_unpin_host": (ctypes.c_int, [_VP, _VP]),
    "st_staged_upload_bytes": (ctypes.c_uint64, [_VP]),
    "st_generate_hilbert": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]),
    "st_generate_uniform": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.c_uint64]),
    "st_solve_device": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.POINTER(StOptions), _VP,
                                       ctypes.POINTER(StResult)]),
    "st_solve_host": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.POINTER(StOptions), _VP, _VP,
                                     ctypes.POINTER(StResult)]),
    "st_solve_streamed": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.POINTER(StOptions), ctypes.c_size_t,
                                         ctypes.c_uint32, _VP, _VP, ctypes.POINTER(StResult),
                                         ctypes.POINTER(StStreamPlan)]),
    "st_solve_file": (ctypes.c_int, [_VP, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(StOptions),
                                     ctypes.c_size_t, ctypes.c_uint32, _VP, _VP, ctypes.POINTER(StResult),
                                     ctypes.POINTER(StStreamPlan)]),
    "st_group_attach": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int), ctypes.c_uint32, ctypes.c_uint32]),
    "st_group_detach": (ctypes.c_int, [_VP]),
    "st_group_size": (ctypes.c_int, [_VP]),
    "st_convert_f32_to_bf16": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "st_solve_device_bf16": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, ctypes.POINTER(StOptions), _VP,
                                            ctypes.POINTER(StResult)]),
    "st_convert_f32_to_fp8": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_uint32, ctypes.c_uint32]),
    "st_solve_device_fp8": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_uint32, ctypes.POINTER(StOptions), _VP,
                                           ctypes.POINTER(StResult)]),
    "st_round_timestamps": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, c_u32p]),
    "st_phase_timestamps": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32, c_u32p]),
    "st_timer_start": (ctypes.c_int, [_VP]),
    "st_timer_stop": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_float)]),
    "st_sum_across_rows": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_uint32]),
    "st_row_pass_readonly": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_uint32, ctypes.c_uint32,
                                            ctypes.c_uint32]),
    "st_find_max": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_uint32]),
    "st_compute_eigen_vector": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_uint32]),
    "st_initialise_eigen_vector": (ctypes.c_int, [_VP, _VP, ctypes.c_uint32]),
    "st_compute_next_matrix": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_uint32]),
    "st_stop": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_uint32, ctypes.c_float]),
    "st_shard_create": (ctypes.c_int, [_VP, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.POINTER(_VP)]),
    "st_shard_export": (ctypes.c_int, [_VP, _VP]),
    "st_shard_import": (ctypes.c_int, [_VP, _VP]),
    "st_shard_link_local": (ctypes.c_int, [ctypes.POINTER(_VP), ctypes.c_uint32]),
    "st_shard_prepare": (ctypes.c_int, [_VP, ctypes.POINTER(StOptions)]),
    "st_shard_rows": (ctypes.c_int, [_VP, c_u32p, c_u32p]),
    "st_shard_solve": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(StOptions), _VP, ctypes.POINTER(StResult)]),
    "st_shard_solve_bf16": (ctypes.c_int, [_VP, _VP, ctypes.POINTER(StOptions), _VP, ctypes.POINTER(StResult)]),
    "st_shard_solve_fp8": (ctypes.c_int, [_VP, _VP, _VP, ctypes.POINTER(StOptions), _VP, ctypes.POINTER(StResult)]),
    "st_shard_destroy": (None, [_VP]),
}

IPC_HANDLE_BYTES = 64
MAX_WORLD = 8

_lib = None


def so_path() -> str:
    return _build.SO_PATH


def load() -> ctypes.CDLL:
    """dlopen libsimilarity_transform.so (building it first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = so_path()
    if _build.stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc on this box: use the shipped binary if there is one
            if not os.path.exists(path):
                raise RuntimeError(
                    f"{_build.SO_NAME} is missing and could not be built ({exc}); "
                    "there is no CPU fallback for the similarity_transform path") from exc
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift, fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class StError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().st_last_error()
        raise StError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
