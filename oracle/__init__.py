"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (eigen_value_b200/) never imports it.

The oracle restates reference similarity_transform.cpp:5-460 on the CPU; see the header of
oracle.c for what pins it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

EPS = 1e-3        # reference include/similarity_transform.hpp:4
MAX_ITR = 1000    # reference include/similarity_transform.hpp:5

FORM_INPLACE, FORM_READONLY = 0, 1
SUM_SEQUENTIAL, SUM_LANES16, SUM_SUBGROUP32 = 0, 1, 2
SUM_CUDA = 4   # the CUDA kernels' evaluation order: bit-identical to the GPU (tests/test_zz_gpu_bitexact.py)
SUM_CUDA_F64 = 6    # SUM_CUDA's order with fp64 accumulators (st_options.accumulate = ST_ACC_F64)
SUM_CUDA_BF16 = SUM_CUDA  # bf16 storage reduces 4-element words in the fp32 kernels' order (feed to_bf16(mat)[0])
SUM_CUDA_FP8 = SUM_CUDA   # fp8 storage reduces 4-element words in the fp32 kernels' order (feed to_fp8_rows(mat)[0])
STOP_ABSOLUTE, STOP_RELATIVE = 0, 1   # the reference's stop test | extension: threshold eps * max(s)


def sum_workgroup(wg_size: int) -> int:
    """The reference's literal two-level summation order for a given work-group size."""
    return 3 | (int(wg_size) << 8)

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so with the recipe in oracle/Makefile."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_sum_across_rows.argtypes = [_f32p, _f32p, ctypes.c_uint32, ctypes.c_int]
        L.oracle_find_max.argtypes = [_f32p, ctypes.c_uint32]
        L.oracle_find_max.restype = ctypes.c_float
        L.oracle_compute_eigen_vector.argtypes = [_f32p, ctypes.c_float, _f32p, ctypes.c_uint32]
        L.oracle_initialise_eigen_vector.argtypes = [_f32p, ctypes.c_uint32]
        L.oracle_stop.argtypes = [_f32p, ctypes.c_uint32, ctypes.c_float]
        L.oracle_stop.restype = ctypes.c_uint32
        L.oracle_compute_next_matrix.argtypes = [_f32p, _f32p, ctypes.c_uint32]
        L.oracle_generate_hilbert.argtypes = [_f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.oracle_generate_uniform.argtypes = [_f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                              ctypes.c_uint64]
        L.oracle_philox_block.argtypes = [ctypes.c_uint64, ctypes.c_uint64,
                                          np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
        L.oracle_similarity_transform_ex.argtypes = [
            _f32p, _f32p, _f32p, ctypes.c_uint32,
            np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"),
            ctypes.c_float, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
            ctypes.POINTER(ctypes.c_double)]
        L.oracle_similarity_transform_ex.restype = ctypes.c_int64
        L.oracle_similarity_transform_ex2.argtypes = [
            _f32p, _f32p, _f32p, ctypes.c_uint32,
            np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"),
            ctypes.c_float, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int,
            ctypes.POINTER(ctypes.c_double)]
        L.oracle_similarity_transform_ex2.restype = ctypes.c_int64
        L.oracle_similarity_transform_generated.argtypes = [
            ctypes.c_int, ctypes.c_uint64, _f32p, _f32p, ctypes.c_uint32,
            np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"),
            ctypes.c_float, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        L.oracle_similarity_transform_generated.restype = ctypes.c_int64
        L.oracle_stop_relative.argtypes = [_f32p, ctypes.c_uint32, ctypes.c_float, ctypes.c_float]
        L.oracle_stop_relative.restype = ctypes.c_uint32
        L.oracle_time_rounds.argtypes = [_f32p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.oracle_time_rounds.restype = ctypes.c_double
        _lib = L
    return _lib


def threads() -> int:
    return int(lib().oracle_threads())


def hilbert(dim: int, row0: int = 0, rows: int | None = None) -> np.ndarray:
    rows = dim - row0 if rows is None else rows
    out = np.empty((rows, dim), dtype=np.float32)
    lib().oracle_generate_hilbert(out, dim, row0, rows)
    return out


def uniform(dim: int, seed: int, row0: int = 0, rows: int | None = None) -> np.ndarray:
    rows = dim - row0 if rows is None else rows
    out = np.empty((rows, dim), dtype=np.float32)
    lib().oracle_generate_uniform(out, dim, row0, rows, seed)
    return out


def to_bf16(mat: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 -> bfloat16, round to nearest even (what cvt.rn.bf16.f32 / st_convert_f32_to_bf16 does; NaN ->
    0x7fff).  Returns (the rounded values widened back to fp32 -- exact --, the uint16 bit patterns).
    Works in row blocks with 32-bit arithmetic so that a 40960^2 matrix costs its two outputs and little more."""
    a = np.ascontiguousarray(mat, dtype=np.float32)
    flat = a.reshape(-1)
    bits = np.empty(flat.shape, dtype=np.uint16)
    back = np.empty(flat.shape, dtype=np.float32)
    step = 1 << 24
    for i in range(0, flat.shape[0], step):
        u = flat[i:i + step].view(np.uint32)
        nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
        # for every non-NaN value u + 0x8000 cannot wrap 32 bits; NaNs are patched afterwards
        r = ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)).astype(np.uint16)
        r[nan] = 0x7FFF
        bits[i:i + step] = r
        back[i:i + step] = (r.astype(np.uint32) << np.uint32(16)).view(np.float32)
    return back.reshape(a.shape), bits.reshape(a.shape)


def fp8_e4m3_values() -> np.ndarray:
    """Value of every e4m3 code 0x00..0x7e (the fn variant: no infinities, 0x7f = NaN), ascending."""
    c = np.arange(127)
    e, m = c >> 3, c & 7
    return np.where(e == 0, np.ldexp(m.astype(np.float64), -9), np.ldexp(1.0 + m / 8.0, e - 7)).astype(np.float32)


def to_fp8_rows(mat: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """fp32 -> fp8 (e4m3) storage with one power-of-two scale per row: what st_convert_f32_to_fp8 does.
    Per row: a = max |x|, scale = 2^k with a / 2^k in (224, 448] (k clamped to +-118; 1 for an all-zero or
    non-finite row); every element divided by the scale (exact) and rounded to the nearest code, ties to the
    even code, saturating at 448 (cvt.rn.satfinite.e4m3x2.f32; NaN -> 0x7f, and the row's scale becomes NaN).
    Returns (the dequantised matrix scale[r] * q[r][c] as fp32 -- exact --, the uint8 codes, the fp32 scales)."""
    a = np.ascontiguousarray(mat, dtype=np.float32)
    rows = a.shape[0]
    vals = fp8_e4m3_values()
    amax = np.max(np.abs(np.where(np.isnan(a), np.float32(0), a)), axis=1)
    m, e = np.frexp(amax)                                    # amax = m * 2^e, m in [0.5, 1)
    k = np.clip(np.where(m <= np.float32(0.875), e - 9, e - 8), -118, 118)
    scale = np.where((amax > 0) & np.isfinite(amax), np.ldexp(np.float32(1), k), np.float32(1)).astype(np.float32)
    has_nan = np.isnan(a).any(axis=1)
    codes = np.empty(a.shape, dtype=np.uint8)
    back = np.empty(a.shape, dtype=np.float32)
    for r in range(rows):
        x = a[r] / scale[r]                                    # exact: a power of two
        mag = np.minimum(np.abs(x), np.float32(448))
        lo = np.clip(np.searchsorted(vals, mag, side="right") - 1, 0, 126)     # largest code <= |x|
        hi = np.minimum(lo + 1, 126)
        dl, dh = mag - vals[lo], vals[hi] - mag
        up = (hi > lo) & ((dh < dl) | ((dh == dl) & ((lo & 1) == 1)))
        c = np.where(up, hi, lo).astype(np.uint8)
        nan = np.isnan(x)
        c[nan] = 0x7F
        sign = np.signbit(x) & ~nan
        codes[r] = c | (sign.astype(np.uint8) << 7)
        v = np.where(nan, np.float32(np.nan), vals[np.minimum(c, 126)])
        back[r] = np.where(sign, -v, v) * scale[r]
    # the kernels read a code's 7 magnitude bits as a number (0x7f would be 480), so a NaN travels in the row's scale:
    # the whole row sum is NaN in every round, as it is in an fp32 solve of a matrix with a NaN in that row
    scale[has_nan] = np.nan
    back[has_nan] = np.nan
    return back, codes, scale


def philox_block(ctr: int, key: int) -> np.ndarray:
    out = np.empty(4, dtype=np.uint32)
    lib().oracle_philox_block(ctr, key, out)
    return out


def sum_across_rows(mat: np.ndarray, sum_mode: int = SUM_LANES16) -> np.ndarray:
    n = mat.shape[0]
    out = np.empty(n, dtype=np.float32)
    lib().oracle_sum_across_rows(np.ascontiguousarray(mat, dtype=np.float32), out, n, sum_mode)
    return out


def find_max(vec: np.ndarray) -> float:
    return float(lib().oracle_find_max(np.ascontiguousarray(vec, dtype=np.float32), vec.shape[0]))


def compute_eigen_vector(vec: np.ndarray, mx: float, eigen_vec: np.ndarray) -> None:
    lib().oracle_compute_eigen_vector(vec, mx, eigen_vec, vec.shape[0])


def stop(vec: np.ndarray, eps: float = EPS) -> int:
    return int(lib().oracle_stop(np.ascontiguousarray(vec, dtype=np.float32), vec.shape[0], eps))


def stop_relative(vec: np.ndarray, eps: float = EPS) -> int:
    """Extension (not reference behaviour): all circular adjacent |diff| < eps * max(0, max(vec))."""
    v = np.ascontiguousarray(vec, dtype=np.float32)
    return int(lib().oracle_stop_relative(v, v.shape[0], eps, find_max(v)))


def compute_next_matrix(mat: np.ndarray, vec: np.ndarray) -> None:
    assert mat.flags["C_CONTIGUOUS"] and mat.dtype == np.float32
    lib().oracle_compute_next_matrix(mat, vec, vec.shape[0])


def similarity_transform(mat: np.ndarray, eps: float = EPS, max_itr: int = MAX_ITR,
                         form: int = FORM_INPLACE, sum_mode: int = SUM_LANES16,
                         ranks: int = 1, stop: int = STOP_ABSOLUTE) -> Tuple[np.float32, np.ndarray, float, int]:
    """(lambda, raw eigen_vec, loop ms, iter_count) -- same tuple as the reference's Python
    wrapper returns (wrapper/python/similarity_transform.py:78)."""
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    n = mat.shape[0]
    assert mat.shape == (n, n)
    val = np.empty(1, dtype=np.float32)
    vec = np.empty(n, dtype=np.float32)
    itr = np.zeros(1, dtype=np.uint32)
    ms = ctypes.c_double(0.0)
    rc = lib().oracle_similarity_transform_ex2(mat, val, vec, n, itr, eps, max_itr, form, sum_mode,
                                               ranks, stop, ctypes.byref(ms))
    if rc < 0:
        raise MemoryError("oracle allocation failed")
    return val[0], vec, ms.value, int(itr[0])


def similarity_transform_generated(kind: str, dim: int, seed: int = 0, eps: float = EPS, max_itr: int = MAX_ITR,
                                   sum_mode: int = SUM_CUDA, stop: int = STOP_ABSOLUTE):
    """The read-only loop on a Hilbert ("hilbert") or seeded uniform ("uniform") matrix that is generated row by
    row and never stored: O(N) memory, the same bits as similarity_transform(hilbert(dim) / uniform(dim, seed),
    form=FORM_READONLY, ...).  How the sizes no host can hold get CPU-computed expected values."""
    val = np.empty(1, dtype=np.float32)
    vec = np.empty(dim, dtype=np.float32)
    itr = np.zeros(1, dtype=np.uint32)
    ms = ctypes.c_double(0.0)
    rc = lib().oracle_similarity_transform_generated({"hilbert": 0, "uniform": 1}[kind], seed, val, vec, dim, itr,
                                                     eps, max_itr, sum_mode, stop, ctypes.byref(ms))
    if rc < 0:
        raise MemoryError("oracle allocation failed")
    return val[0], vec, ms.value, int(itr[0])


def time_rounds(mat: np.ndarray, rounds: int, form: int = FORM_INPLACE) -> float:
    """ms for `rounds` full reference rounds (no early exit) on all host threads."""
    n = mat.shape[0]
    return float(lib().oracle_time_rounds(mat, n, rounds, form))
