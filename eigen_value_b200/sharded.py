"""Row-block sharded similarity_transform(): one process per GPU.

`torch.distributed` is plumbing only: it ships each rank's 64-byte CUDA IPC handle to the
peers once (and, in the collective variant, all-gathers the row-sum slices).  The product
path is `ShardedSolver.solve`: a single persistent kernel per GPU whose round barrier doubles
as the exchange -- every CTA stores its row sums straight into all peers' buffers over
NVLink and one flag per peer closes the round (csrc/kernels.cuh: round_barrier).

The reference has no multi-device code at all; the partition is the one
BASELINE.json names: rank g owns rows [dim*g/world, dim*(g+1)/world).
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import StResult, check
from .similarity_transform import (EPS, MAX_ITR, DeviceBuffer, SolveInfo, Solver, make_options)


def shard_rows(dim: int, rank: int, world: int) -> Tuple[int, int]:
    """(row0, rows) of rank's block; same arithmetic as st_shard_create (csrc/abi.cu)."""
    if not (0 <= rank < world <= dim):
        raise ValueError("need 0 <= rank < world <= dim")
    r0 = dim * rank // world
    r1 = dim * (rank + 1) // world
    return r0, r1 - r0


def exchange_handles(handle: bytes, rank: int, world: int, group=None) -> List[bytes]:
    """All-gather of fixed-size opaque handles over torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist
    if len(handle) != _lib.IPC_HANDLE_BYTES:
        raise ValueError("handle must be 64 bytes")
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    table = torch.empty(world * _lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(table, mine, group=group)
    raw = bytes(table.cpu().tolist())
    out = [raw[i * _lib.IPC_HANDLE_BYTES:(i + 1) * _lib.IPC_HANDLE_BYTES] for i in range(world)]
    if out[rank] != handle:
        raise RuntimeError("handle table is not in rank order")
    return out


class ShardedSolver:
    """This rank's shard of a dim x dim problem plus the mapped exchange blocks of its peers."""

    def __init__(self, solver: Solver, dim: int, rank: int, world: int,
                 exchange: Optional[Callable[[bytes, int, int], Sequence[bytes]]] = None,
                 barrier: Optional[Callable[[], None]] = None):
        self.solver, self.dim, self.rank, self.world = solver, dim, rank, world
        self._barrier = barrier
        self._prepared = (1000, 0)      # (max_iter, in-place form) st_shard_create reserved scratch for
        self._vec: Optional[DeviceBuffer] = None
        self.row0, self.rows = shard_rows(dim, rank, world)
        self.lib = solver.lib
        self.shard = ctypes.c_void_p()
        check(self.lib.st_shard_create(solver.ctx, dim, rank, world, ctypes.byref(self.shard)), "st_shard_create")
        r0, rows = ctypes.c_uint32(), ctypes.c_uint32()
        check(self.lib.st_shard_rows(self.shard, ctypes.byref(r0), ctypes.byref(rows)), "st_shard_rows")
        assert (r0.value, rows.value) == (self.row0, self.rows)
        if world > 1:
            buf = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            check(self.lib.st_shard_export(self.shard, buf), "st_shard_export")
            table = (exchange or exchange_handles)(buf.raw, rank, world)
            blob = b"".join(table)
            assert len(blob) == world * _lib.IPC_HANDLE_BYTES
            check(self.lib.st_shard_import(self.shard, blob), "st_shard_import")

    def hilbert(self) -> DeviceBuffer:
        return self.solver.hilbert(self.dim, self.row0, self.rows)

    def uniform(self, seed: int) -> DeviceBuffer:
        return self.solver.uniform(self.dim, seed, self.row0, self.rows)

    def solve(self, d_rows: DeviceBuffer, d_eigen_vec: Optional[DeviceBuffer] = None,
              bf16: bool = False, fp8_scale: Optional[DeviceBuffer] = None, **opts) -> Tuple[SolveInfo, Optional[np.ndarray]]:
        """Collective: every rank calls it with its own rows.  Every rank gets the full
        eigenvector and identical (lambda, iter_count).  The eigenvector is downloaded only
        when no device output buffer is supplied."""
        o = make_options(self.lib, **opts)
        res = StResult()
        own = d_eigen_vec is None
        # A sharded solve never allocates on the device (an allocation may wait for peers that already spin in the
        # collective kernel): scratch for options beyond what st_shard_create reserved is taken here, on every rank
        # (the options of a collective call are the same everywhere), followed by a host barrier.
        need = (int(o.max_iter), 1 if int(o.form) != 0 else 0)
        grow = need[0] > self._prepared[0] or need[1] > self._prepared[1] or (own and self._vec is None)
        if grow:
            check(self.lib.st_shard_prepare(self.shard, ctypes.byref(o)), "st_shard_prepare")
            if own and self._vec is None:
                self._vec = self.solver.alloc(4 * self.dim)
            self._prepared = (max(need[0], self._prepared[0]), max(need[1], self._prepared[1]))
            self.solver.synchronize()
            self._host_barrier()
        vec = self._vec if own else d_eigen_vec
        if fp8_scale is not None:                                                 # rows in fp8 storage + their row scales
            check(self.lib.st_shard_solve_fp8(self.shard, d_rows.ptr, fp8_scale.ptr, ctypes.byref(o), vec.ptr,
                                              ctypes.byref(res)), "st_shard_solve_fp8")
        else:
            fn = self.lib.st_shard_solve_bf16 if bf16 else self.lib.st_shard_solve   # bf16: rows in bfloat16 storage
            check(fn(self.shard, d_rows.ptr, ctypes.byref(o), vec.ptr, ctypes.byref(res)), "st_shard_solve")
        out = vec.download(np.float32, self.dim) if own else None
        return SolveInfo.from_c(res), out

    def _host_barrier(self) -> None:
        if self.world == 1:
            return
        if self._barrier is not None:
            self._barrier()
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            return
        raise RuntimeError("ShardedSolver: scratch for these options had to be reserved (st_shard_prepare), which must be "
                           "followed by a host barrier across the ranks; pass barrier=... when torch.distributed is not used")

    def close(self) -> None:
        if self._vec is not None:
            self._vec.free()
            self._vec = None
        if self.shard is not None and self.shard.value:
            self.lib.st_shard_destroy(self.shard)
            self.shard = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# collective variant: host-driven rounds, torch.distributed all-gather of the row-sum slices
# ---------------------------------------------------------------------------------------------
class RoundBackend:
    """The four per-round operations of reference similarity_transform.cpp:40-44 on one rank's
    row block.  `CudaRoundBackend` runs them as CUDA kernels through the C ABI on torch CUDA
    tensors; the CPU tests inject a numpy implementation to exercise the loop under gloo."""

    def row_pass(self, e, s_slice) -> None:          # s_slice <- (A_g . e) / e_g
        raise NotImplementedError

    def find_max(self, s) -> float:
        raise NotImplementedError

    def stop(self, s, eps: float) -> bool:
        raise NotImplementedError

    def update(self, s, m: float, e) -> None:        # e *= s / m
        raise NotImplementedError


class CudaRoundBackend(RoundBackend):
    def __init__(self, solver: Solver, d_rows: DeviceBuffer, dim: int, row0: int, rows: int):
        import torch
        self.torch = torch
        self.solver, self.d_rows, self.dim, self.row0, self.rows = solver, d_rows, dim, row0, rows
        dev = torch.device("cuda", solver.device)
        self._m = torch.zeros(1, dtype=torch.float32, device=dev)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._full = torch.zeros(dim, dtype=torch.float32, device=dev)

    def _sync_in(self):
        self.torch.cuda.current_stream().synchronize()   # torch stream -> solver stream hand-off

    def row_pass(self, e, s_slice) -> None:
        self._sync_in()
        lib, ctx = self.solver.lib, self.solver.ctx
        check(lib.st_row_pass_readonly(ctx, self.d_rows.ptr, e.data_ptr(), self._full.data_ptr(), self.dim,
                                       self.row0, self.rows), "st_row_pass_readonly")
        self.solver.synchronize()
        s_slice.copy_(self._full[self.row0:self.row0 + self.rows])

    def find_max(self, s) -> float:
        self._sync_in()
        check(self.solver.lib.st_find_max(self.solver.ctx, s.data_ptr(), self._m.data_ptr(), self.dim), "st_find_max")
        self.solver.synchronize()
        return float(self._m.item())

    def stop(self, s, eps: float) -> bool:
        self._sync_in()
        check(self.solver.lib.st_stop(self.solver.ctx, s.data_ptr(), self._flag.data_ptr(), self.dim, eps), "st_stop")
        self.solver.synchronize()
        return bool(self._flag.item())

    def update(self, s, m: float, e) -> None:
        self._sync_in()
        self._m.fill_(m)
        self._sync_in()
        check(self.solver.lib.st_compute_eigen_vector(self.solver.ctx, s.data_ptr(), self._m.data_ptr(),
                                                      e.data_ptr(), self.dim), "st_compute_eigen_vector")
        self.solver.synchronize()


def collective_round_loop(backend: RoundBackend, dim: int, rank: int, world: int, device="cpu",
                          eps: float = EPS, max_iter: int = MAX_ITR, group=None):
    """Read-only-form round loop with a torch.distributed all-gather per round (the baseline
    the fused peer-store kernel is measured against).  Uneven blocks are padded to the
    largest block for the gather.  Returns (lambda, e, iter_count)."""
    import torch
    import torch.distributed as dist
    row0, rows = shard_rows(dim, rank, world)
    blocks = [shard_rows(dim, g, world) for g in range(world)]
    pad = max(b[1] for b in blocks)
    e = torch.ones(dim, dtype=torch.float32, device=device)
    s = torch.empty(dim, dtype=torch.float32, device=device)
    mine = torch.zeros(pad, dtype=torch.float32, device=device)
    gathered = torch.empty(world * pad, dtype=torch.float32, device=device)
    it = max_iter
    for i in range(max_iter):
        backend.row_pass(e, mine[:rows])
        if world > 1:
            dist.all_gather_into_tensor(gathered, mine, group=group)
            for g, (r0, n) in enumerate(blocks):
                s[r0:r0 + n] = gathered[g * pad:g * pad + n]
        else:
            s[row0:row0 + rows] = mine[:rows]
        m = backend.find_max(s)
        backend.update(s, m, e)
        if backend.stop(s, eps):
            it = i
            break
    return float(s[0].item()), e, it
