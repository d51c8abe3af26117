"""Host-side mirror of the reference's Python interface, on top of the CUDA C ABI.

`EigenValue` has the reference's name, constructor and `similarity_transform(mat)` method
with the same argument checks and the same 4-tuple result
(reference wrapper/python/similarity_transform.py:18-78).  `Solver` is the additive
surface: device-resident matrices, on-device Hilbert / uniform generation, solver options,
per-kernel calls (reference include/similarity_transform.hpp:55-100) and per-round timing.

No numerics happen in this file: every call goes through libsimilarity_transform.so.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import StOptions, StResult, StStreamPlan, check

EPS = 1e-3       # reference include/similarity_transform.hpp:4
MAX_ITR = 1000   # reference include/similarity_transform.hpp:5
FORM_READONLY, FORM_INPLACE = 0, 1
STOP_ABSOLUTE, STOP_RELATIVE = 0, 1   # st_options.stop: the reference's test | max diff < eps * max(s)
ACC_F32, ACC_F64 = 0, 1               # st_options.accumulate: fp32 like the reference | fp64 accumulators


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class EigenValue:
    """Drop-in for the reference's `EigenValue` (similarity_transform.py:18-78)."""

    def __init__(self, devices=None, min_dim: int = 0) -> None:
        # reference :29-40 -- load the shared object, make the queue, fail on a NULL handle
        self.so_lib = _lib.load()
        self.so_path = _lib.so_path()
        self.sycl_q = ctypes.c_void_p()          # keeps the reference's attribute name
        self.so_lib.make_queue(ctypes.byref(self.sycl_q))
        if self.sycl_q.value is None:
            err = self.so_lib.st_last_error()
            raise Exception("failed to get default CUDA device queue: "
                            + (err.decode() if err else "no GPU"))
        # extension (st_group_attach): devices="all" or a list of helper GPUs -- matrices of min_dim rows and
        # up (0 = 8192) are then row-block sharded over all of them behind the same similarity_transform() call.
        # The unmodified reference wrapper gets the same from the ST_DEVICES environment variable.
        if devices is not None:
            attach_group(self.so_lib, self.sycl_q, devices, min_dim)

    @property
    def device_count(self) -> int:
        return int(self.so_lib.st_group_size(self.sycl_q))

    def pinned(self, mat: np.ndarray):
        """Context manager (extension): page-locks `mat` for the duration of the block, so that repeated
        similarity_transform(mat) calls move it at full PCIe rate instead of through the driver's staging
        buffers (st_pin_host / st_unpin_host).

            with ev.pinned(mat):
                lam, vec, ms, rounds = ev.similarity_transform(mat)
        """
        import contextlib

        @contextlib.contextmanager
        def scope():
            assert mat.flags["C_CONTIGUOUS"]
            check(self.so_lib.st_pin_host(self.sycl_q, _ptr(mat), mat.nbytes), "st_pin_host")
            try:
                yield mat
            finally:
                check(self.so_lib.st_unpin_host(self.sycl_q, _ptr(mat)), "st_unpin_host")
        return scope()

    def similarity_transform(self, mat: np.ndarray) -> Tuple[np.float32, np.ndarray, int, int]:
        """(max eigen value, eigen vector, loop milliseconds, iteration count) of a positive
        square float32 matrix; A v = lambda v holds to the reference's tolerance
        (wrapper/python/test.py:15-16)."""
        m, n = mat.shape
        assert m == n, "must be square matrix of floating points !"          # reference :55
        assert mat.dtype.num == 11, "dtype of input matrix must be float32 !"  # reference :56-57
        assert mat.flags["C_CONTIGUOUS"], "matrix must be C-contiguous"       # ndpointer flag, :59-60

        eigen_val = np.empty(1, dtype=np.float32)
        eigen_vec = np.empty(n, dtype=np.float32)
        iter_cnt = np.zeros(1, dtype=np.uint)    # 8-byte slot, C writes the low 4 bytes (:73)
        ts = self.so_lib.max_eigen_value(self.sycl_q, _ptr(mat), _ptr(eigen_val), _ptr(eigen_vec),
                                         n, _ptr(iter_cnt))
        if ts < 0:
            err = self.so_lib.st_last_error()
            raise RuntimeError(f"max_eigen_value failed ({ts}): {err.decode() if err else ''}")
        return eigen_val[0], eigen_vec, ts, int(iter_cnt[0])


def attach_group(lib, ctx, devices="all", min_dim: int = 0) -> None:
    """st_group_attach: bind helper GPUs ("all" = every other visible one, or a list of device ids that does
    not name the context's own device) to a context / make_queue handle."""
    if isinstance(devices, str):
        assert devices == "all", devices
        check(lib.st_group_attach(ctx, None, 0, min_dim), "st_group_attach")
    else:
        ids = (ctypes.c_int * len(devices))(*devices)
        if len(devices) == 0:
            return
        check(lib.st_group_attach(ctx, ids, len(devices), min_dim), "st_group_attach")


@dataclass
class SolveInfo:
    eigen_val: np.float32
    iter_count: int
    passes: int
    launches: int
    loop_ms: float
    total_ms: float
    round_us_median: float
    round_us_min: float
    bytes_per_round: int
    grid: int
    kernel_id: int = 0
    threads: int = 0

    KERNEL_NAMES = {1: "st::round_loop_kernel", 2: "st::round_loop_wide_kernel", 10: "st::round_loop_sc_kernel",
                    20: "st::round_loop_cluster_kernel",
                    30: "st::sum_across_rows_kernel + tail kernels (streamed, host-driven rounds)"}

    @property
    def kernel_name(self) -> str:
        k = self.kernel_id
        key = 30 if k >= 30 else 20 if k >= 20 else 10 if k >= 10 else 2 if k == 2 else 1
        return self.KERNEL_NAMES[key]

    @classmethod
    def from_c(cls, r: StResult) -> "SolveInfo":
        return cls(np.float32(r.eigen_val), int(r.iter_count), int(r.passes), int(r.launches),
                   float(r.loop_ms), float(r.total_ms), float(r.round_us_median),
                   float(r.round_us_min), int(r.bytes_per_round), int(r.grid), int(r.kernel_id),
                   int(r.threads))


class DeviceBuffer:
    """A device allocation owned by a Solver (plain pointer + size; no torch types)."""

    def __init__(self, solver: "Solver", nbytes: int):
        self.solver = solver
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(solver.lib.st_malloc(solver.ctx, self.nbytes, ctypes.byref(p)), "st_malloc")
        self.ptr = p

    def free(self) -> None:
        if self.ptr is not None and self.ptr.value:
            self.solver.lib.st_free(self.solver.ctx, self.ptr)
            self.ptr = None

    def upload(self, a: np.ndarray) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(self.solver.lib.st_memcpy_h2d(self.solver.ctx, self.ptr, _ptr(a), a.nbytes), "st_memcpy_h2d")
        return self

    def download(self, dtype, count: int) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(self.solver.lib.st_memcpy_d2h(self.solver.ctx, _ptr(out), self.ptr, out.nbytes), "st_memcpy_d2h")
        return out

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_options(lib, eps: float = EPS, max_iter: int = MAX_ITR, form: int = FORM_READONLY,
                 sweep: Optional[int] = None, threads: int = 0, ctas: int = 0, kernel: int = 0,
                 l2_keep_pct: Optional[int] = None, stop: int = STOP_ABSOLUTE,
                 accumulate: int = ACC_F32) -> StOptions:
    o = StOptions()
    lib.st_default_options(ctypes.byref(o))
    o.eps, o.max_iter, o.form, o.stop, o.accumulate = eps, max_iter, form, stop, accumulate
    o.threads, o.ctas, o.kernel = threads, ctas, kernel
    if sweep is not None:
        o.sweep = sweep
    if l2_keep_pct is not None:
        o.l2_keep_pct = l2_keep_pct
    return o


class Solver:
    """One CUDA device: st_create / st_destroy plus everything that runs on it."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        self.ctx = ctypes.c_void_p()
        check(self.lib.st_create(device, ctypes.byref(self.ctx)), "st_create")
        self.device = device
        sm = ctypes.c_int()
        l2 = ctypes.c_size_t()
        hbm = ctypes.c_size_t()
        name = ctypes.create_string_buffer(128)
        check(self.lib.st_device_info(self.ctx, ctypes.byref(sm), ctypes.byref(l2), ctypes.byref(hbm),
                                      name, 128), "st_device_info")
        self.sm_count, self.l2_bytes, self.hbm_bytes = sm.value, l2.value, hbm.value
        self.name = name.value.decode()

    def close(self) -> None:
        if self.ctx is not None and self.ctx.value:
            self.lib.st_destroy(self.ctx)
            self.ctx = None

    def attach_group(self, devices="all", min_dim: int = 0) -> int:
        """Helper GPUs behind this solver (st_group_attach): solve_host() then shards matrices of min_dim rows
        and up (0 = 8192) over all of them.  Returns the number of devices now behind the handle."""
        attach_group(self.lib, self.ctx, devices, min_dim)
        return int(self.lib.st_group_size(self.ctx))

    def detach_group(self) -> None:
        check(self.lib.st_group_detach(self.ctx), "st_group_detach")

    # ---- memory / inputs ------------------------------------------------------------------
    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def upload(self, a: np.ndarray) -> DeviceBuffer:
        a = np.ascontiguousarray(a)
        return self.alloc(max(a.nbytes, 4)).upload(a)

    def hilbert(self, dim: int, row0: int = 0, rows: Optional[int] = None) -> DeviceBuffer:
        """reference utils.cpp:136-154, generated on the device."""
        rows = dim - row0 if rows is None else rows
        buf = self.alloc(4 * rows * dim)
        check(self.lib.st_generate_hilbert(self.ctx, buf.ptr, dim, row0, rows), "st_generate_hilbert")
        return buf

    def uniform(self, dim: int, seed: int, row0: int = 0, rows: Optional[int] = None) -> DeviceBuffer:
        """seeded uniform (0,1] fill (replaces reference utils.cpp:124-134)."""
        rows = dim - row0 if rows is None else rows
        buf = self.alloc(4 * rows * dim)
        check(self.lib.st_generate_uniform(self.ctx, buf.ptr, dim, row0, rows, seed), "st_generate_uniform")
        return buf

    def to_bf16(self, d_src: DeviceBuffer, count: int) -> DeviceBuffer:
        """fp32 -> bfloat16 (round to nearest even) on the device: storage for solve_device(..., bf16=True)."""
        out = self.alloc(2 * count)
        check(self.lib.st_convert_f32_to_bf16(self.ctx, d_src.ptr, out.ptr, count), "st_convert_f32_to_bf16")
        return out

    def to_fp8(self, d_src: DeviceBuffer, rows: int, dim: int) -> Tuple[DeviceBuffer, DeviceBuffer]:
        """fp32 -> fp8 (e4m3) storage with one power-of-two scale per row, on the device: (codes, row scales), the
        storage for solve_device(..., fp8_scale=...).  dim % 4 == 0."""
        codes, scale = self.alloc(rows * dim), self.alloc(4 * rows)
        check(self.lib.st_convert_f32_to_fp8(self.ctx, d_src.ptr, codes.ptr, scale.ptr, rows, dim), "st_convert_f32_to_fp8")
        return codes, scale

    def synchronize(self) -> None:
        check(self.lib.st_synchronize(self.ctx), "st_synchronize")

    # ---- the round loop -------------------------------------------------------------------
    def solve_device(self, d_mat: DeviceBuffer, dim: int, d_eigen_vec: Optional[DeviceBuffer] = None,
                     bf16: bool = False, fp8_scale: Optional[DeviceBuffer] = None,
                     **opts) -> Tuple[SolveInfo, Optional[np.ndarray]]:
        """similarity_transform() on a device-resident matrix.  Returns (info, eigen_vec);
        eigen_vec is downloaded only when no device output buffer was supplied.  bf16=True: d_mat
        holds bfloat16 storage (see to_bf16; opt-in, changes results, dim % 4 == 0).  fp8_scale=<row scales>: d_mat
        holds fp8 storage (see to_fp8; opt-in, changes results, dim % 4 == 0)."""
        o = make_options(self.lib, **opts)
        res = StResult()
        own = d_eigen_vec is None
        vec = self.alloc(4 * dim) if own else d_eigen_vec
        if fp8_scale is not None:
            assert not bf16
            check(self.lib.st_solve_device_fp8(self.ctx, d_mat.ptr, fp8_scale.ptr, dim, ctypes.byref(o), vec.ptr,
                                               ctypes.byref(res)), "st_solve_device_fp8")
        else:
            fn, what = (self.lib.st_solve_device_bf16, "st_solve_device_bf16") if bf16 else \
                       (self.lib.st_solve_device, "st_solve_device")
            check(fn(self.ctx, d_mat.ptr, dim, ctypes.byref(o), vec.ptr, ctypes.byref(res)), what)
        out = vec.download(np.float32, dim) if own else None
        if own:
            vec.free()
        return SolveInfo.from_c(res), out

    def solve_host(self, mat: np.ndarray, **opts) -> Tuple[SolveInfo, np.ndarray]:
        """similarity_transform() on a host matrix (H2D copy, solve, D2H of the results)."""
        assert mat.ndim == 2 and mat.shape[0] == mat.shape[1] and mat.dtype == np.float32
        assert mat.flags["C_CONTIGUOUS"]
        n = mat.shape[0]
        o = make_options(self.lib, **opts)
        res = StResult()
        val = np.empty(1, dtype=np.float32)
        vec = np.empty(n, dtype=np.float32)
        check(self.lib.st_solve_host(self.ctx, _ptr(mat), n, ctypes.byref(o), _ptr(val), _ptr(vec),
                                     ctypes.byref(res)), "st_solve_host")
        return SolveInfo.from_c(res), vec

    def solve_streamed(self, mat, device_budget: int = 0, block_rows: int = 0, offset: int = 0,
                       dim: Optional[int] = None, **opts) -> Tuple[SolveInfo, np.ndarray, dict]:
        """similarity_transform() on a host matrix that does not fit the device (or the `device_budget`
        bytes granted): a device cache of row blocks + alternating sweeps, only the uncached blocks cross
        PCIe per round (st_solve_streamed).  `mat` is a C-contiguous float32 array -- np.memmap works --
        or a file path (then `dim`, and `offset` = bytes before the first element; st_solve_file).
        Same bits as solve_host.  Returns (info, eigen_vec, plan)."""
        o = make_options(self.lib, **opts)
        res, plan = StResult(), StStreamPlan()
        val = np.empty(1, dtype=np.float32)
        if isinstance(mat, (str, bytes)):
            assert dim is not None, "a file needs dim"
            vec = np.empty(dim, dtype=np.float32)
            path = mat.encode() if isinstance(mat, str) else mat
            check(self.lib.st_solve_file(self.ctx, path, offset, dim, ctypes.byref(o), device_budget, block_rows,
                                         _ptr(val), _ptr(vec), ctypes.byref(res), ctypes.byref(plan)), "st_solve_file")
        else:
            assert mat.ndim == 2 and mat.shape[0] == mat.shape[1] and mat.dtype == np.float32
            assert mat.flags["C_CONTIGUOUS"]
            n = mat.shape[0]
            vec = np.empty(n, dtype=np.float32)
            check(self.lib.st_solve_streamed(self.ctx, _ptr(mat), n, ctypes.byref(o), device_budget, block_rows,
                                             _ptr(val), _ptr(vec), ctypes.byref(res), ctypes.byref(plan)),
                  "st_solve_streamed")
        return SolveInfo.from_c(res), vec, {name: int(getattr(plan, name)) for name, _ in StStreamPlan._fields_}

    def solve_tensor(self, mat, **opts):
        """Zero-copy solve of a square float32 CUDA tensor (torch, or anything exposing
        `__cuda_array_interface__`) that lives on this solver's device: no PCIe staging, which is
        what dominates `max_eigen_value` from N = 8192 up.  Returns (info, eigen_vec) with
        eigen_vec a torch tensor on the same device when torch is importable, else numpy."""
        iface = getattr(mat, "__cuda_array_interface__", None)
        if iface is None:
            raise TypeError("solve_tensor needs an object with __cuda_array_interface__")
        shape, typestr, strides = iface["shape"], iface["typestr"], iface.get("strides")
        if len(shape) != 2 or shape[0] != shape[1]:
            raise ValueError("must be square matrix of floating points !")     # reference :55
        if typestr not in ("<f4", "=f4"):
            raise ValueError("dtype of input matrix must be float32 !")        # reference :56-57
        n = int(shape[0])
        if strides is not None and tuple(strides) != (4 * n, 4):
            raise ValueError("matrix must be C-contiguous")
        ptr = ctypes.c_void_p(int(iface["data"][0]))
        try:
            import torch
        except ImportError:  # pragma: no cover
            torch = None
        if torch is not None and isinstance(mat, torch.Tensor):
            if mat.device.index is not None and mat.device.index != self.device:
                raise ValueError("tensor lives on another device than this solver")
            torch.cuda.current_stream(mat.device).synchronize()   # hand-off to the solver's stream
            out = torch.empty(n, dtype=torch.float32, device=mat.device)
            out_ptr = ctypes.c_void_p(out.data_ptr())
        else:
            out, out_ptr = None, None
        o = make_options(self.lib, **opts)
        res = StResult()
        if out is not None:
            check(self.lib.st_solve_device(self.ctx, ptr, n, ctypes.byref(o), out_ptr, ctypes.byref(res)),
                  "st_solve_device")
            return SolveInfo.from_c(res), out
        vec = self.alloc(4 * n)
        check(self.lib.st_solve_device(self.ctx, ptr, n, ctypes.byref(o), vec.ptr, ctypes.byref(res)),
              "st_solve_device")
        host = vec.download(np.float32, n)
        vec.free()
        return SolveInfo.from_c(res), host

    def round_timestamps(self) -> np.ndarray:
        n = ctypes.c_uint32()
        check(self.lib.st_round_timestamps(self.ctx, None, 0, ctypes.byref(n)), "st_round_timestamps")
        out = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            check(self.lib.st_round_timestamps(self.ctx, _ptr(out), n.value, ctypes.byref(n)),
                  "st_round_timestamps")
        return out

    def phase_breakdown(self) -> dict:
        """Median microseconds per round of the last solve, split at CTA 0's stamps:
        matrix pass | barrier (+ exchange) | vector tail."""
        n = ctypes.c_uint32()
        check(self.lib.st_phase_timestamps(self.ctx, None, 0, ctypes.byref(n)), "st_phase_timestamps")
        ph = np.zeros(n.value, dtype=np.uint64)
        if n.value:
            check(self.lib.st_phase_timestamps(self.ctx, _ptr(ph), n.value, ctypes.byref(n)),
                  "st_phase_timestamps")
        rounds = self.round_timestamps().astype(np.int64)
        ph = ph.astype(np.int64).reshape(-1, 3)
        if len(ph) == 0:
            return {}
        start = rounds[:-1]
        med = lambda a: float(np.median(a)) * 1e-3
        return {"pass_us": med(ph[:, 0] - start), "barrier_us": med(ph[:, 1] - ph[:, 0]),
                "tail_us": med(ph[:, 2] - ph[:, 1]), "rounds": int(len(ph))}

    def timer_start(self) -> None:
        """CUDA-event stopwatch on the solver's stream (st_timer_start)."""
        check(self.lib.st_timer_start(self.ctx), "st_timer_start")

    def timer_stop(self) -> float:
        """Milliseconds since timer_start(), measured on the device (st_timer_stop; synchronises)."""
        ms = ctypes.c_float()
        check(self.lib.st_timer_stop(self.ctx, ctypes.byref(ms)), "st_timer_stop")
        return float(ms.value)

    # ---- per-kernel calls (reference L1 functions), numpy in / numpy out -------------------
    def sum_across_rows(self, mat: np.ndarray) -> np.ndarray:
        n = mat.shape[0]
        d_mat, d_vec = self.upload(mat.astype(np.float32)), self.alloc(4 * n)
        check(self.lib.st_sum_across_rows(self.ctx, d_mat.ptr, d_vec.ptr, n), "st_sum_across_rows")
        self.synchronize()
        return d_vec.download(np.float32, n)

    def find_max(self, vec: np.ndarray) -> np.float32:
        n = vec.shape[0]
        d_vec, d_max = self.upload(vec.astype(np.float32)), self.alloc(4)
        check(self.lib.st_find_max(self.ctx, d_vec.ptr, d_max.ptr, n), "st_find_max")
        self.synchronize()
        return d_max.download(np.float32, 1)[0]

    def compute_eigen_vector(self, vec: np.ndarray, mx: float, eigen_vec: np.ndarray) -> np.ndarray:
        n = vec.shape[0]
        d_vec, d_max = self.upload(vec.astype(np.float32)), self.upload(np.array([mx], dtype=np.float32))
        d_e = self.upload(eigen_vec.astype(np.float32))
        check(self.lib.st_compute_eigen_vector(self.ctx, d_vec.ptr, d_max.ptr, d_e.ptr, n),
              "st_compute_eigen_vector")
        self.synchronize()
        return d_e.download(np.float32, n)

    def initialise_eigen_vector(self, n: int) -> np.ndarray:
        d_e = self.alloc(4 * n)
        check(self.lib.st_initialise_eigen_vector(self.ctx, d_e.ptr, n), "st_initialise_eigen_vector")
        self.synchronize()
        return d_e.download(np.float32, n)

    def compute_next_matrix(self, mat: np.ndarray, vec: np.ndarray) -> np.ndarray:
        n = mat.shape[0]
        d_mat, d_vec = self.upload(mat.astype(np.float32)), self.upload(vec.astype(np.float32))
        check(self.lib.st_compute_next_matrix(self.ctx, d_mat.ptr, d_vec.ptr, n), "st_compute_next_matrix")
        self.synchronize()
        return d_mat.download(np.float32, n * n).reshape(n, n)

    def stop(self, vec: np.ndarray, eps: float = EPS) -> int:
        n = vec.shape[0]
        d_vec, d_ret = self.upload(vec.astype(np.float32)), self.alloc(4)
        check(self.lib.st_stop(self.ctx, d_vec.ptr, d_ret.ptr, n, eps), "st_stop")
        self.synchronize()
        return int(d_ret.download(np.uint32, 1)[0])

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
