"""Streamed solve (host / file-backed input for matrices larger than the device).

st_solve_streamed keeps a direct-mapped device cache of row blocks, sweeps the blocks in alternating
direction and drives the rounds from the host (include/similarity_transform.h).  Its contract: the bits
of the in-device solve (hence of the oracle in the CUDA kernels' summation order) whatever the block
size and the cache size are, and exactly (blocks - slots) blocks over PCIe per round after the first.

STATUS: written after round 1's GPU budget was spent; runs on the emulated library inside the CPU
suite (tests/test_gpu_suite_emulated.py).  Sorts last so that it cannot disturb the tests before it.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from eigen_value_b200 import FORM_INPLACE, STOP_RELATIVE
from eigen_value_b200._lib import StResult, StStreamPlan

pytestmark = pytest.mark.gpu


def _want(mat, **kw):
    return oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, **kw)


def _same_bits(info, vec, want):
    w_val, w_vec, _, w_it = want
    assert info.iter_count == w_it, (info.iter_count, w_it)
    assert np.float32(info.eigen_val).view(np.uint32) == np.float32(w_val).view(np.uint32), (float(info.eigen_val), float(w_val))
    assert np.array_equal(vec.view(np.uint32), w_vec.view(np.uint32))


def _matrix(kind, dim):
    if kind == "hilbert":
        return oracle.hilbert(dim)
    return (oracle.uniform(dim, 3000 + dim) + np.float32(0.25)).astype(np.float32)


# (dim, block_rows, slots): one block; everything cached; two slots (pure double buffering); a short last
# block; a cache that wraps (blocks % slots != 0); scalar rows (dim % 4 != 0); rows of two 8192-column chunks
CASES = [(3, 3, 1), (64, 16, 4), (64, 16, 2), (1000, 96, 3), (1000, 96, 7), (1024, 100, 4), (1023, 64, 5),
         (8200, 1024, 3)]


@pytest.mark.parametrize("kind", ["hilbert", "uniform"])
@pytest.mark.parametrize("dim,block_rows,slots", CASES)
def test_streamed_solve_has_the_bits_of_the_in_device_solve(solver, kind, dim, block_rows, slots):
    mat = _matrix(kind, dim)
    keep = mat.copy()
    cap = 12 if (kind == "uniform" and dim > 4100) else 1000
    budget = slots * block_rows * dim * 4
    info, vec, plan = solver.solve_streamed(mat, device_budget=budget, block_rows=block_rows, max_iter=cap)
    assert np.array_equal(mat, keep)                                  # the caller's matrix is never modified
    _same_bits(info, vec, _want(mat, max_itr=cap))
    blocks = -(-dim // block_rows)
    assert plan["streamed"] == 1 and plan["blocks"] == blocks and plan["slots"] == min(slots, blocks)
    assert plan["block_rows"] == min(block_rows, dim)
    assert info.kernel_id == 30 and info.passes == min(info.iter_count + 1, cap)
    # PCIe traffic: the whole matrix once, then only what the alternating sweep does not find in the cache
    assert plan["h2d_bytes_first"] == 4 * dim * dim
    if info.passes > 1:
        last_rows = dim - (blocks - 1) * block_rows
        missed = blocks - plan["slots"]
        if missed == 0:
            assert plan["h2d_bytes_per_round"] == 0
        else:
            full = 4 * dim * block_rows
            # the short last block is among the misses of a backward round only if it is not cached then
            assert plan["h2d_bytes_per_round"] in (missed * full, (missed - 1) * full + 4 * dim * last_rows)
        assert plan["h2d_bytes_total"] <= 4 * dim * dim + (info.passes - 1) * missed * 4 * dim * block_rows


def test_streamed_solve_with_the_relative_stop_test(solver):
    # a matrix the reference's absolute test cannot finish early keeps streaming until the relative test holds
    dim = 1000
    mat = _matrix("uniform", dim)
    info, vec, plan = solver.solve_streamed(mat, device_budget=3 * 128 * dim * 4, block_rows=128, eps=1e-6,
                                            stop=STOP_RELATIVE, max_iter=60)
    _same_bits(info, vec, _want(mat, eps=1e-6, stop=oracle.STOP_RELATIVE, max_itr=60))
    assert 0 < info.iter_count < 60


def test_streamed_solve_runs_to_the_cap_on_nan_and_the_handle_survives(solver):
    dim = 64
    mat = _matrix("uniform", dim)
    mat[5, 7] = np.nan
    info, vec, plan = solver.solve_streamed(mat, device_budget=2 * 16 * dim * 4, block_rows=16, max_iter=25)
    assert info.iter_count == 25 and info.passes == 25
    good = oracle.hilbert(256)
    info, vec, _ = solver.solve_streamed(good, device_budget=2 * 64 * 256 * 4, block_rows=64)
    assert info.iter_count == 10                                      # reference README.md:71
    _same_bits(info, vec, _want(good))


def test_automatic_budget_hands_a_matrix_that_fits_to_the_fused_solve(solver):
    mat = oracle.hilbert(512)
    info, vec, plan = solver.solve_streamed(mat)
    assert plan["streamed"] == 0 and info.kernel_id != 30 and info.launches == 1
    _same_bits(info, vec, _want(mat))


def test_file_backed_matrix(solver, tmp_path):
    # a raw dump and a .npy file (offset = its header): st_solve_file maps the file and streams it
    dim = 1000
    mat = _matrix("hilbert", dim)
    want = _want(mat)
    raw = tmp_path / "hilbert.f32"
    mat.tofile(raw)
    info, vec, plan = solver.solve_streamed(str(raw), dim=dim, device_budget=4 * 64 * dim * 4, block_rows=64)
    _same_bits(info, vec, want)
    assert plan["streamed"] == 1 and plan["blocks"] == 16 and plan["slots"] == 4
    npy = tmp_path / "hilbert.npy"
    np.save(npy, mat)
    header = os.path.getsize(npy) - mat.nbytes
    info, vec, _ = solver.solve_streamed(str(npy), dim=dim, offset=header, device_budget=3 * 100 * dim * 4, block_rows=100)
    _same_bits(info, vec, want)
    # np.memmap goes through the host-pointer entry point
    mm = np.load(npy, mmap_mode="r")
    info, vec, _ = solver.solve_streamed(mm, device_budget=3 * 100 * dim * 4, block_rows=100)
    _same_bits(info, vec, want)


def test_bad_arguments_are_refused_cleanly(solver, tmp_path):
    lib = solver.lib
    mat = oracle.hilbert(64)
    res, plan = StResult(), StStreamPlan()
    val = np.empty(1, np.float32)
    vec = np.empty(64, np.float32)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)

    def streamed(h, dim, budget, rows, opt=None):
        return lib.st_solve_streamed(solver.ctx, h, dim, opt, budget, rows, p(val), p(vec), ctypes.byref(res), ctypes.byref(plan))

    assert streamed(None, 64, 0, 0) == -2
    assert streamed(p(mat), 0, 0, 0) == -2
    assert streamed(p(mat), 64, 16 * 64 * 4, 16) == -2 and b"two row blocks" in lib.st_last_error()   # one slot only
    from eigen_value_b200.similarity_transform import make_options
    o = make_options(lib, form=FORM_INPLACE)
    assert streamed(p(mat), 64, 4 * 16 * 64 * 4, 16, ctypes.byref(o)) == -2            # read-only form only
    o = make_options(lib, accumulate=1)
    assert streamed(p(mat), 64, 4 * 16 * 64 * 4, 16, ctypes.byref(o)) == -2            # fp32 accumulation only

    def from_file(path, offset, dim):
        return lib.st_solve_file(solver.ctx, path, offset, dim, None, 4 * 16 * 64 * 4, 16, p(val), p(vec),
                                 ctypes.byref(res), ctypes.byref(plan))

    assert from_file(None, 0, 64) == -2
    assert from_file(str(tmp_path / "missing.f32").encode(), 0, 64) == -2 and b"cannot open" in lib.st_last_error()
    short = tmp_path / "short.f32"
    mat[:10].tofile(short)
    assert from_file(str(short).encode(), 0, 64) == -2 and b"shorter" in lib.st_last_error()
    whole = tmp_path / "whole.f32"
    mat.tofile(whole)
    assert from_file(str(whole).encode(), 2, 64) == -2                                  # offset not a multiple of 4
    assert from_file(str(whole).encode(), 0, 64) == 0 and res.iter_count == _want(mat)[3]
    # the context is fine afterwards
    info, _ = solver.solve_device(solver.hilbert(256), 256)
    assert info.iter_count == 10


def test_random_block_and_cache_shapes_never_change_a_bit(solver):
    # 80 random (dim, block_rows, slots) plans, both sweep parities of every cache geometry: always the in-device bits,
    # always (blocks - slots) blocks per later round, or a clean refusal when the budget holds one block of several
    rng = np.random.default_rng(20261018)
    for _ in range(80):
        dim = int(rng.integers(1, 200))
        block_rows = int(rng.integers(1, dim + 1))
        blocks = -(-dim // block_rows)
        slots = int(rng.integers(1, blocks + 2))
        mat = (rng.random((dim, dim), dtype=np.float32) + np.float32(0.05)).astype(np.float32)
        budget = slots * block_rows * dim * 4 + int(rng.integers(0, 4 * dim))          # slack below one more block
        if slots < 2 and blocks > 1:
            with pytest.raises(Exception, match="two row blocks"):
                solver.solve_streamed(mat, device_budget=budget, block_rows=block_rows)
            continue
        cap = int(rng.integers(1, 9))
        info, vec, plan = solver.solve_streamed(mat, device_budget=budget, block_rows=block_rows, max_iter=cap)
        base, base_vec = solver.solve_host(mat, max_iter=cap)
        assert info.iter_count == base.iter_count and info.passes == base.passes, (dim, block_rows, slots)
        assert np.float32(info.eigen_val).view(np.uint32) == np.float32(base.eigen_val).view(np.uint32)
        assert np.array_equal(vec.view(np.uint32), base_vec.view(np.uint32)), (dim, block_rows, slots)
        assert plan["blocks"] == blocks and plan["slots"] == min(slots, blocks)
        rows_missed = plan["h2d_bytes_per_round"] // (4 * dim)
        if info.passes > 1:
            missed = blocks - plan["slots"]
            last_rows = dim - (blocks - 1) * block_rows
            assert rows_missed in (missed * block_rows, max(0, missed - 1) * block_rows + (last_rows if missed else 0))


def test_automatic_block_size_halves_until_two_blocks_fit_the_budget(solver):
    # block_rows = 0: about 64 MiB per block, halved until the budget holds two (launch_plan.hpp: stream_shape_for)
    dim = 1000
    mat = _matrix("hilbert", dim)
    info, vec, plan = solver.solve_streamed(mat, device_budget=1_000_000)       # 250 rows of 4000 bytes
    _same_bits(info, vec, _want(mat))
    assert plan["block_rows"] == 125 and plan["blocks"] == 8 and plan["slots"] == 2 and plan["cache_bytes"] == 2 * 125 * 4000
    # a budget below two rows cannot double-buffer anything
    with pytest.raises(Exception, match="two row blocks"):
        solver.solve_streamed(mat, device_budget=7000)
