"""ctypes front-end of oracle/_ref/libreference_cpu.so: the reference's UNMODIFIED C++ sources
(similarity_transform.cpp, utils.cpp, wrapper/similarity_transform.cpp) compiled against the
single-threaded CPU SYCL shim in oracle/sycl_shim (recipe: `make -C oracle ref`).

TEST INFRASTRUCTURE ONLY.  It exists in the build container (where /root/reference is mounted)
and travels to the GPU box as a built binary; `available()` says whether it is there.  It is
used to validate oracle.c bit for bit and to generate tests/golden/reference_sycl.json.
"""
from __future__ import annotations

import ctypes
import os
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libreference_cpu.so")
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_lib = None


def available() -> bool:
    return os.path.exists(SO)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(SO)
        u = ctypes.c_uint
        L.ref_max_eigen_value.argtypes = [_f32, _f32, _f32, u, _u32]
        L.ref_max_eigen_value.restype = ctypes.c_int64
        L.ref_similarity_transform.argtypes = [_f32, _f32, _f32, u, u, _u32]
        L.ref_similarity_transform.restype = ctypes.c_int64
        L.ref_sum_across_rows.argtypes = [_f32, _f32, u, u]
        L.ref_find_max.argtypes = [_f32, _f32, u, u]
        L.ref_compute_eigen_vector.argtypes = [_f32, _f32, _f32, u, u]
        L.ref_initialise_eigen_vector.argtypes = [_f32, u]
        L.ref_compute_next_matrix.argtypes = [_f32, _f32, u, u]
        L.ref_stop.argtypes = [_f32, _u32, u, u]
        L.ref_generate_hilbert_matrix.argtypes = [_f32, u]
        L.ref_identity_matrix.argtypes = [_f32, u, u]
        L.ref_generate_vector.argtypes = [_f32, u, u]
        L.ref_stop_criteria_test_success_data.argtypes = [_f32, u, u]
        L.ref_stop_criteria_test_fail_data.argtypes = [_f32, u, u]
        L.ref_max_work_group_size.restype = u
        _lib = L
    return _lib


def wrapper_wg_size(dim: int) -> int:
    """reference wrapper/similarity_transform.cpp:33"""
    return min(dim >> 1, int(lib().ref_max_work_group_size()))


def max_eigen_value(mat: np.ndarray) -> Tuple[np.float32, np.ndarray, int, int]:
    """make_queue + max_eigen_value exactly as the reference's Python wrapper drives them."""
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    n = mat.shape[0]
    val, vec, it = np.empty(1, np.float32), np.empty(n, np.float32), np.zeros(1, np.uint32)
    keep = mat.copy()
    ms = lib().ref_max_eigen_value(mat, val, vec, n, it)
    if ms < 0:
        raise RuntimeError("reference rejected the launch shape (dim % wg_size != 0)")
    assert np.array_equal(mat, keep)
    return val[0], vec, int(ms), int(it[0])


def similarity_transform(mat: np.ndarray, wg_size: int) -> Tuple[np.float32, np.ndarray, int, int]:
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    n = mat.shape[0]
    val, vec, it = np.empty(1, np.float32), np.empty(n, np.float32), np.zeros(1, np.uint32)
    ms = lib().ref_similarity_transform(mat, val, vec, n, wg_size, it)
    if ms < 0:
        raise RuntimeError("reference rejected the launch shape")
    return val[0], vec, int(ms), int(it[0])


def hilbert(dim: int) -> np.ndarray:
    out = np.empty((dim, dim), np.float32)
    assert lib().ref_generate_hilbert_matrix(out, dim) == 0
    return out
