"""ctypes front-end of the CPU emulation harness (tests/cuda_emu/libcuda_emu.so): the round kernels of
eigen_value_b200/csrc, compiled for the host, run with small launch shapes.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as _build  # noqa: E402


class Opts(ctypes.Structure):
    _fields_ = [("eps", ctypes.c_float), ("max_iter", ctypes.c_uint32), ("form", ctypes.c_int32),
                ("sweep", ctypes.c_int32), ("dynamic", ctypes.c_int32), ("threads", ctypes.c_int32),
                ("ctas", ctypes.c_int32), ("kernel", ctypes.c_int32), ("stop", ctypes.c_int32),
                ("bf16", ctypes.c_int32), ("world", ctypes.c_int32), ("acc64", ctypes.c_int32),
                ("row_scale", ctypes.c_void_p)]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        u32p, f32p, vp = ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_float), ctypes.c_void_p
        L.emu_last_error.restype = ctypes.c_char_p
        L.emu_solve.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(Opts), f32p, vp, u32p, u32p, u32p]
        L.emu_find_max.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint, f32p]
        L.emu_stop.argtypes = [vp, ctypes.c_uint32, ctypes.c_float, ctypes.c_uint, u32p]
        L.emu_convert_bf16.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_uint]
        L.emu_convert_fp8.argtypes = [vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint]
        L.emu_sum_across_rows.argtypes = [vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint]
        _lib = L
    return _lib


def solve(mat: np.ndarray, dim: int, kernel: int = 1, threads: int = 64, ctas: int = 3, world: int = 1,
          eps: float = 1e-3, max_iter: int = 1000, form: int = 0, sweep: int = 1, dynamic: int = -1,
          stop: int = 0, bf16: bool = False, acc64: bool = False, fp8_scale=None):
    """One solve on `world` emulated GPUs of `ctas` CTAs x `threads` threads each.
    Returns (lambda, eigen_vec, iter_count, passes, all_ranks_agree).  fp8_scale: the row scales of fp8 storage
    (mat then holds the uint8 codes)."""
    fp8 = fp8_scale is not None
    assert mat.flags["C_CONTIGUOUS"] and mat.dtype == (np.uint8 if fp8 else np.uint16 if bf16 else np.float32)
    if fp8:
        fp8_scale = np.ascontiguousarray(fp8_scale, dtype=np.float32)
    o = Opts(eps=eps, max_iter=max_iter, form=form, sweep=sweep, dynamic=dynamic, threads=threads, ctas=ctas,
             kernel=kernel, stop=stop, bf16=2 if fp8 else int(bf16), world=world, acc64=int(acc64),
             row_scale=fp8_scale.ctypes.data if fp8 else None)
    val, it, ps, agree = ctypes.c_float(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(1)
    vec = np.empty(dim, dtype=np.float32)
    rc = lib().emu_solve(mat.ctypes.data, dim, ctypes.byref(o), ctypes.byref(val), vec.ctypes.data,
                         ctypes.byref(it), ctypes.byref(ps), ctypes.byref(agree))
    if rc != 0:
        raise RuntimeError(lib().emu_last_error().decode())
    return np.float32(val.value), vec, int(it.value), int(ps.value), bool(agree.value)


def find_max(vec: np.ndarray, ctas: int = 3) -> np.float32:
    v = np.ascontiguousarray(vec, dtype=np.float32)
    out = ctypes.c_float()
    lib().emu_find_max(v.ctypes.data, v.shape[0], ctas, ctypes.byref(out))
    return np.float32(out.value)


def stop(vec: np.ndarray, eps: float = 1e-3, ctas: int = 3) -> int:
    v = np.ascontiguousarray(vec, dtype=np.float32)
    out = ctypes.c_uint32()
    lib().emu_stop(v.ctypes.data, v.shape[0], eps, ctas, ctypes.byref(out))
    return int(out.value)


def convert_bf16(x: np.ndarray, ctas: int = 2) -> np.ndarray:
    src = np.ascontiguousarray(x, dtype=np.float32)
    dst = np.empty(src.shape, dtype=np.uint16)
    lib().emu_convert_bf16(src.ctypes.data, dst.ctypes.data, src.size, ctas)
    return dst


def convert_fp8(x: np.ndarray, ctas: int = 2):
    """(codes, row scales) of st_convert_f32_to_fp8's kernel."""
    src = np.ascontiguousarray(x, dtype=np.float32)
    codes = np.empty(src.shape, dtype=np.uint8)
    scale = np.empty(src.shape[0], dtype=np.float32)
    lib().emu_convert_fp8(src.ctypes.data, codes.ctypes.data, scale.ctypes.data, src.shape[0], src.shape[1], ctas)
    return codes, scale


def sum_across_rows(mat: np.ndarray, e=None, row0: int = 0, ctas: int = 3) -> np.ndarray:
    m = np.ascontiguousarray(mat, dtype=np.float32)
    rows, dim = m.shape
    vec = np.zeros(dim, dtype=np.float32)
    ep = np.ascontiguousarray(e, dtype=np.float32) if e is not None else None
    lib().emu_sum_across_rows(m.ctypes.data, ep.ctypes.data if ep is not None else None, vec.ctypes.data, dim, row0, rows, ctas)
    return vec[row0:row0 + rows]
