"""The CUDA round kernels' LOGIC, executed on the CPU.

tests/cuda_emu compiles the kernel sources of eigen_value_b200/csrc for the host (every CUDA thread a
fiber, every CTA an OS thread, every emulated GPU a group of them; see cuda_emu.h) and runs them with
small launch shapes.  The results must equal the oracle evaluated in the kernels' summation order
(oracle.SUM_CUDA / SUM_CUDA_BF16) BIT FOR BIT: eigenvalue, raw eigenvector, round count.  This covers,
without a GPU, the index arithmetic, the reduction order, the work-unit scheduling (static, dynamic,
rows of several units), the prefetch bookkeeping, both stop tests, bf16 storage, the cluster kernel's
distributed-shared-memory exchange and -- with several emulated GPUs running concurrently -- the
fused row-sum exchange and its flag barrier.  It says nothing about the hardware: memory-model
subtleties, alignment faults and speed are what the `-m gpu` suite and the profiles are for.

Nothing here is the product: the product path needs the CUDA library and a GPU.
"""
import os
import sys

import numpy as np
import pytest

import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import emu  # noqa: E402

A3 = np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)     # reference tests/test.cpp:84-94


def U(dim):
    return (oracle.uniform(dim, seed=1000 + dim) + np.float32(0.25)).astype(np.float32)


def expect(mat, form=0, sum_mode=oracle.SUM_CUDA, **kw):
    o_form = oracle.FORM_INPLACE if form else oracle.FORM_READONLY
    val, vec, _, it = oracle.similarity_transform(mat, form=o_form, sum_mode=sum_mode, eps=kw.get("eps", 1e-3),
                                                  max_itr=kw.get("max_iter", 1000), stop=kw.get("stop", 0))
    return val, vec, it


def same_bits(got, want):
    val, vec, it, passes, agree = got
    w_val, w_vec, w_it = want
    assert agree, "emulated GPUs returned different results"
    assert it == w_it, (it, w_it)
    assert np.float32(val).view(np.uint32) == np.float32(w_val).view(np.uint32), (float(val), float(w_val))
    assert np.array_equal(vec.view(np.uint32), w_vec.view(np.uint32))


# ---- general chunked loop (round_loop_kernel) ------------------------------------------------------
@pytest.mark.parametrize("dim,threads,ctas", [(1, 32, 1), (2, 32, 1), (3, 32, 1), (5, 32, 2), (31, 64, 2), (100, 64, 3),
                                              (257, 96, 4), (640, 64, 5), (1000, 128, 3), (1023, 64, 4)])
@pytest.mark.parametrize("form", [0, 1])
def test_general_loop(dim, threads, ctas, form):
    mat = A3 if dim == 3 else U(dim)
    same_bits(emu.solve(mat, dim, kernel=1, threads=threads, ctas=ctas, form=form), expect(mat, form))


def test_three_by_three_golden_and_hilbert_round_counts():
    val, vec, it, _, _ = emu.solve(A3, 3, kernel=1, threads=32, ctas=1)
    assert it == 4 and abs(val - 7.53114) < 1e-3                        # reference tests/test.cpp:96-102
    for dim, rounds in ((128, 9), (256, 10), (512, 12), (1024, 13)):      # reference README.md:70-73
        got = emu.solve(oracle.hilbert(dim), dim, kernel=13, threads=128, ctas=4)
        assert got[2] == rounds
        same_bits(got, expect(oracle.hilbert(dim)))


@pytest.mark.parametrize("sweep", [0, 1])
@pytest.mark.parametrize("threads,ctas", [(32, 1), (64, 7), (256, 2)])
def test_launch_shape_and_sweep_do_not_change_bits(sweep, threads, ctas):
    mat = U(520)
    want = expect(mat)
    for kernel in (1, 13):
        same_bits(emu.solve(mat, 520, kernel=kernel, threads=threads, ctas=ctas, sweep=sweep), want)


# ---- resident-e kernel (round_loop_sc_kernel): prefetch slots, static / dynamic units ------------------
@pytest.mark.parametrize("kernel", [10, 11, 12, 13])
@pytest.mark.parametrize("dynamic", [0, 1])
def test_resident_e_kernel_configurations(kernel, dynamic):
    for dim, threads, ctas in ((8, 32, 1), (640, 64, 3), (1000, 64, 5)):
        mat = U(dim)
        same_bits(emu.solve(mat, dim, kernel=kernel, threads=threads, ctas=ctas, dynamic=dynamic), expect(mat))


def test_resident_rows_stay_in_the_prefetch_slot():
    # whole rows fit the slot and no warp owns two units: fetched once, served from shared memory every round
    mat = oracle.hilbert(2048)
    same_bits(emu.solve(mat, 2048, kernel=10, threads=512, ctas=4), expect(mat))   # 64 warps >= ... not resident
    mat = oracle.hilbert(64)
    same_bits(emu.solve(mat, 64, kernel=13, threads=128, ctas=16), expect(mat))     # 64 warps, 64 one-unit rows: resident
    # the shapes the B200 launch plan produces (512 threads; one row per warp): N = 128 on 8 CTAs, N = 1024 on 64 --
    # lanes that reach the mbarrier wait before lane 0 has issued the copy must not starve it
    for dim, ctas, kernel in ((128, 8, 13), (1024, 64, 13), (2048, 128, 10)):
        mat = oracle.hilbert(dim)
        same_bits(emu.solve(mat, dim, kernel=kernel, threads=512, ctas=ctas), expect(mat))


@pytest.mark.parametrize("kernel,dynamic", [(13, 1), (13, 0), (11, 1), (1, -1)])
def test_rows_of_several_work_units(kernel, dynamic):
    # N = 8200: every row is two units (8192 + 8 columns); chunk sums combined by whoever completes the row
    mat = U(8200)
    same_bits(emu.solve(mat, 8200, kernel=kernel, threads=128, ctas=6, dynamic=dynamic, max_iter=4),
              expect(mat, max_iter=4))


# ---- resident-e kernel on scalar units: dim % 4 != 0 (configuration 11, 4-byte loads, scalar vector tail) -----
@pytest.mark.parametrize("dim,threads,ctas,dynamic", [(1, 32, 1, 0), (2, 32, 1, 1), (3, 32, 1, 0), (5, 32, 2, 1), (31, 64, 2, 0),
                                                      (257, 96, 4, 1), (1001, 128, 3, 0), (1023, 64, 4, 1), (1030, 512, 2, 1)])
def test_resident_e_kernel_on_scalar_units(dim, threads, ctas, dynamic):
    mat = A3 if dim == 3 else U(dim)
    want = expect(mat)
    same_bits(emu.solve(mat, dim, kernel=11, threads=threads, ctas=ctas, dynamic=dynamic), want)
    same_bits(emu.solve(mat, dim, kernel=1, threads=threads, ctas=ctas), want)       # the general loop's order, bit for bit


def test_resident_e_kernel_on_scalar_units_extras():
    mat = U(8195)                                                            # two units per row: 8192 + 3 columns
    same_bits(emu.solve(mat, 8195, kernel=11, threads=128, ctas=6, dynamic=1, max_iter=3), expect(mat, max_iter=3))
    same_bits(emu.solve(mat, 8195, kernel=11, threads=128, ctas=3, world=2, max_iter=3), expect(mat, max_iter=3))
    mat = U(2051)                                                            # 2048 + 3 columns
    same_bits(emu.solve(mat, 2051, kernel=11, threads=64, ctas=5, dynamic=1, max_iter=5), expect(mat, max_iter=5))
    mat = U(1025)
    same_bits(emu.solve(mat, 1025, kernel=11, threads=64, ctas=5, max_iter=5), expect(mat, max_iter=5))
    mat = U(1001)                                                            # sharded, ranks of 333 / 334 / 334 rows
    same_bits(emu.solve(mat, 1001, kernel=11, threads=64, ctas=3, world=3), expect(mat))
    kw = dict(stop=1, eps=1e-6, max_iter=40)                                 # relative stop test
    same_bits(emu.solve(mat, 1001, kernel=11, threads=64, ctas=2, **kw), expect(mat, **kw))
    H = oracle.hilbert(1022)                                                 # a dozen rounds, alternating sweep on and off
    for sweep in (0, 1):
        same_bits(emu.solve(H, 1022, kernel=11, threads=96, ctas=5, sweep=sweep), expect(H))


# ---- on-chip cluster kernel ---------------------------------------------------------
@pytest.mark.parametrize("dim", [4, 8, 100, 128, 384, 512])
def test_cluster_kernel_distributed_shared_memory(dim):
    mat = oracle.hilbert(dim) if dim >= 128 else U(dim)
    same_bits(emu.solve(mat, dim, kernel=20), expect(mat))


# ---- stop tests ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel,dim,form", [(1, 1000, 0), (1, 1023, 1), (1, 1023, 0), (13, 1000, 0), (10, 640, 0), (20, 512, 0), (20, 100, 0)])
def test_relative_stop(kernel, dim, form):
    mat = U(dim)
    for eps in (1e-3, 1e-6):
        kw = dict(eps=eps, stop=1, max_iter=60)
        same_bits(emu.solve(mat, dim, kernel=kernel, threads=64, ctas=4, form=form, **kw), expect(mat, form, **kw))


def test_nan_runs_to_the_cap_under_both_stop_tests():
    mat = U(64)
    mat[32, 21] = np.nan
    for stop in (0, 1):
        for kernel in (1, 13, 20):
            assert emu.solve(mat, 64, kernel=kernel, threads=64, ctas=2, stop=stop, max_iter=30)[2] == 30


def test_eps_and_cap():
    mat = oracle.hilbert(256)
    for kw in (dict(eps=1e-2), dict(max_iter=5), dict(eps=0.0, max_iter=25)):
        same_bits(emu.solve(mat, 256, kernel=13, threads=64, ctas=4, **kw), expect(mat, **kw))


# ---- bf16 storage ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [1, 11])
@pytest.mark.parametrize("dim", [8, 64, 1000, 2048])
def test_bf16_storage(kernel, dim):
    rounded, bits = oracle.to_bf16(U(dim))
    for kw in (dict(), dict(stop=1, eps=1e-6, max_iter=60)):
        same_bits(emu.solve(bits, dim, kernel=kernel, threads=64, ctas=4, bf16=True, **kw),
                  expect(rounded, sum_mode=oracle.SUM_CUDA_BF16, **kw))


@pytest.mark.parametrize("kernel", [1, 11])
def test_bf16_storage_rows_of_several_units(kernel):
    rounded, bits = oracle.to_bf16(U(8200))
    same_bits(emu.solve(bits, 8200, kernel=kernel, threads=128, ctas=6, bf16=True, max_iter=3),
              expect(rounded, sum_mode=oracle.SUM_CUDA_BF16, max_iter=3))


def test_bf16_conversion_kernel():
    for n in (1, 7, 8, 1000, 4096 * 5 + 3):
        x = (np.random.default_rng(n).random(n) * 100 + 1e-3).astype(np.float32)
        x[: min(n, 4)] = np.array([1.00390625, 1.01171875, 1.0, 3.0e38], dtype=np.float32)[: min(n, 4)]
        assert np.array_equal(emu.convert_bf16(x), oracle.to_bf16(x)[1])
    assert emu.convert_bf16(np.array([np.nan], np.float32))[0] == 0x7FFF
    # misaligned source: the scalar path
    base = np.arange(1, 102, dtype=np.float32) / 7
    assert np.array_equal(emu.convert_bf16(base[1:]), oracle.to_bf16(base[1:])[1])


# ---- fp8 (e4m3) storage with one power-of-two scale per row --------------------------------------------------
def test_fp8_conversion_kernel():
    rng = np.random.default_rng(8)
    x = (rng.random((37, 64)) * np.exp(rng.normal(0, 6, (37, 1)))).astype(np.float32)     # rows of very different magnitude
    x[3] = 0                                                                              # an all-zero row keeps scale 1
    x[5, :4] = [448.0, -448.0, 1e-9, -3.0]                                                # saturation edge, underflow, sign
    x[7, 9] = np.nan
    x[11] *= np.float32(1e30)
    x[12] *= np.float32(1e-30)
    back, codes, scale = oracle.to_fp8_rows(x)
    got_codes, got_scale = emu.convert_fp8(x)
    assert np.array_equal(got_scale, scale, equal_nan=True) and np.isnan(scale[7]) and np.array_equal(got_codes, codes)
    fin = np.isfinite(x) & ~np.isnan(scale)[:, None]
    assert np.all(np.abs(back - x)[fin] <= np.abs(x)[fin] / 16 + (scale[:, None] * np.float32(2.0 ** -10) * np.ones_like(x))[fin])
    # every code value survives the round trip: 0x00..0x7e and their negatives, in one row with the largest at 448
    vals = oracle.fp8_e4m3_values()
    row = np.concatenate([vals, -vals, np.zeros(2, np.float32)])[None, :].astype(np.float32)
    b2, c2, s2 = oracle.to_fp8_rows(row)
    assert s2[0] == 1 and np.array_equal(b2, row) and np.array_equal(emu.convert_fp8(row)[0], c2)


@pytest.mark.parametrize("kernel,dim,threads,ctas,world", [(11, 4, 32, 1, 1), (11, 16, 32, 1, 1), (11, 636, 64, 3, 1), (11, 1008, 128, 4, 2),
                                                           (1, 644, 64, 3, 1), (1, 1008, 96, 2, 3)])
def test_fp8_storage(kernel, dim, threads, ctas, world):
    back, codes, scale = oracle.to_fp8_rows(U(dim))
    same_bits(emu.solve(codes, dim, kernel=kernel, threads=threads, ctas=ctas, world=world, fp8_scale=scale),
              expect(back, sum_mode=oracle.SUM_CUDA_FP8))


def test_fp8_storage_extras():
    back, codes, scale = oracle.to_fp8_rows(U(8196))                         # two units per row: 8192 + 4 columns
    for kernel in (11, 1):
        same_bits(emu.solve(codes, 8196, kernel=kernel, threads=128, ctas=6, max_iter=3, fp8_scale=scale),
                  expect(back, sum_mode=oracle.SUM_CUDA_FP8, max_iter=3))
    H = oracle.hilbert(512)                                                  # Hilbert: entries down to 1/1023 of the row's first
    back, codes, scale = oracle.to_fp8_rows(H)
    kw = dict(stop=1, eps=1e-5, max_iter=60)
    same_bits(emu.solve(codes, 512, kernel=11, threads=64, ctas=4, fp8_scale=scale, **kw),
              expect(back, sum_mode=oracle.SUM_CUDA_FP8, **kw))
    val = emu.solve(codes, 512, kernel=11, threads=64, ctas=4, fp8_scale=scale)[0]
    assert abs(val - expect(H)[0]) < 0.02 * expect(H)[0]                     # the storage format moves lambda by ~1 %


# ---- several emulated GPUs: the fused exchange and its flat barrier ---------------------------------------
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kernel", [1, 13])
def test_row_block_sharding_is_bit_identical_to_one_gpu(world, kernel):
    for dim, form in ((1000, 0), (640, 0)) + (((1001, 1),) if kernel == 1 else ()):
        mat = U(dim)
        same_bits(emu.solve(mat, dim, kernel=kernel, threads=64, ctas=2, world=world, form=form), expect(mat, form))


def test_flat_barrier_with_unequal_grids():
    # 33 rows on 2 GPUs: 16 and 17 rows, one warp per CTA -> the ranks run 16 and 17 CTAs; each GPU still adds
    # exactly kArriveUnits per round.  Hilbert: a dozen rounds, so the three max slots go round several times.
    H = oracle.hilbert(36)
    same_bits(emu.solve(H, 36, kernel=13, threads=32, ctas=64, world=2), expect(H))
    same_bits(emu.solve(H, 36, kernel=1, threads=32, ctas=64, world=3), expect(H))


def test_sharded_extras():
    mat = U(8200)                                                            # two units per row, two GPUs
    same_bits(emu.solve(mat, 8200, kernel=13, threads=128, ctas=3, world=2, max_iter=3), expect(mat, max_iter=3))
    rounded, bits = oracle.to_bf16(U(1000))                                   # bf16 storage + relative stop, four GPUs
    kw = dict(stop=1, eps=1e-6, max_iter=40)
    same_bits(emu.solve(bits, 1000, kernel=11, threads=64, ctas=2, world=4, bf16=True, **kw),
              expect(rounded, sum_mode=oracle.SUM_CUDA_BF16, **kw))
    H = oracle.hilbert(512)                                                   # 12 rounds of barrier crossings on 8 GPUs
    same_bits(emu.solve(H, 512, kernel=13, threads=32, ctas=1, world=8), expect(H))


# ---- standalone kernels --------------------------------------------------------------------------------------
def test_standalone_find_max_and_stop():
    N = 1 << 14
    assert emu.find_max(np.arange(1, N + 1, dtype=np.float32)) == N            # reference tests/test.cpp:32-41
    assert emu.find_max(-np.ones(5, dtype=np.float32)) == 0.0                  # zero-filled cell, reference :169
    assert emu.find_max(np.array([1.0, np.nan, 3.0], np.float32)) == 3.0
    ok = np.full(N, np.float32(1.0001), dtype=np.float32)
    bad = (np.arange(1, N + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    assert emu.stop(ok) == oracle.stop(ok) == 1
    assert emu.stop(bad) == oracle.stop(bad) == 0                              # fails through the wrap pair only
    edge = np.array([0.0, 1e-3, 0.0, 0.0], dtype=np.float32)
    assert emu.stop(edge) == oracle.stop(edge) == 0                            # strict <
    assert emu.stop(np.array([5.0], np.float32)) == 1


@pytest.mark.parametrize("dim", [7, 64, 1000, 8200])
def test_standalone_row_sum_kernel(dim):
    m = oracle.uniform(dim, seed=99 + dim)[: min(dim, 40)]
    full = np.zeros((dim, dim), dtype=np.float32)
    full[: m.shape[0]] = m
    assert np.array_equal(emu.sum_across_rows(m), oracle.sum_across_rows(full, oracle.SUM_CUDA)[: m.shape[0]])


# ---- fp64 accumulation (st_options.accumulate = ST_ACC_F64) ------------------------------------------------
@pytest.mark.parametrize("kernel,dim,threads,ctas", [(1, 5, 32, 1), (1, 100, 64, 3), (1, 1023, 64, 3), (13, 1000, 64, 4),
                                                     (10, 640, 64, 3), (12, 2048, 128, 4)])
def test_fp64_accumulation(kernel, dim, threads, ctas):
    mat = U(dim)
    for kw in (dict(), dict(stop=1, eps=1e-6, max_iter=60)):
        same_bits(emu.solve(mat, dim, kernel=kernel, threads=threads, ctas=ctas, acc64=True, **kw),
                  expect(mat, sum_mode=oracle.SUM_CUDA_F64, **kw))


def test_fp64_accumulation_rows_of_several_units_and_sharded():
    mat = U(8200)
    same_bits(emu.solve(mat, 8200, kernel=13, threads=128, ctas=6, acc64=True, max_iter=3),
              expect(mat, sum_mode=oracle.SUM_CUDA_F64, max_iter=3))
    mat = U(1000)
    same_bits(emu.solve(mat, 1000, kernel=13, threads=64, ctas=2, world=3, acc64=True), expect(mat, sum_mode=oracle.SUM_CUDA_F64))



def test_general_loop_with_several_staged_windows_per_row(monkeypatch):
    # The general loop stages up to 32768 columns of the scale vector at a time and reduces a row in 8192-column chunks
    # inside the window.  With the window shrunk to 8192 / 16384 columns a 16400-column row crosses three / two windows
    # (the last one 16 columns wide): same chunk sums, same left-to-right order, same bits.
    mat = U(16400)
    want = expect(mat, max_iter=2)
    for window in ("8192", "16384", "32768"):
        monkeypatch.setenv("ST_EMU_WINDOW", window)
        same_bits(emu.solve(mat, 16400, kernel=1, threads=128, ctas=5, max_iter=2), want)
    monkeypatch.setenv("ST_EMU_WINDOW", "8192")
    same_bits(emu.solve(mat, 16400, kernel=1, threads=64, ctas=3, world=2, form=1, max_iter=2), expect(mat, 1, max_iter=2))


# ---- wide kernel (kernel 2): unit-scheduled, the eigenvector staged one window at a time --------------------------
@pytest.mark.parametrize("dim,threads,ctas,world,dynamic,sweep", [(520, 64, 3, 1, 1, 1), (520, 64, 3, 1, 0, 1), (1000, 128, 4, 1, 1, 0),
                                                                  (36, 32, 2, 1, 1, 1), (4, 32, 1, 1, 0, 1), (1000, 64, 2, 3, 1, 1),
                                                                  (640, 96, 5, 2, 0, 1)])
def test_wide_kernel_one_window(dim, threads, ctas, world, dynamic, sweep):
    mat = U(dim)
    same_bits(emu.solve(mat, dim, kernel=2, threads=threads, ctas=ctas, world=world, dynamic=dynamic, sweep=sweep), expect(mat))


def test_wide_kernel_stop_tests_and_round_cap():
    mat = U(256)
    for kw in (dict(eps=1e-2), dict(max_iter=5), dict(eps=0.0, max_iter=25), dict(eps=1e-6, stop=1, max_iter=60)):
        okw = dict(kw)
        if "stop" in okw:
            okw["stop"] = oracle.STOP_RELATIVE
        same_bits(emu.solve(mat, 256, kernel=2, threads=64, ctas=4, **kw), expect(mat, **okw))
    bad = mat.copy()
    bad[100, 21] = np.nan
    assert emu.solve(bad, 256, kernel=2, threads=64, ctas=2, max_iter=30)[2] == 30


def test_wide_kernel_several_windows_per_row(monkeypatch):
    # 16400 columns = three 8192-column chunks (the last one 16 columns); windows of 8192 columns -> three phases per
    # round (one chunk each), 16384 -> two phases (two chunks + one); every CTA walks them on its own, every window
    # has its own unit counter.  Alternating sweep: odd rounds go through the windows backwards.
    mat = U(16400)
    want = expect(mat, max_iter=3)
    for window, dyn in (("8192", 1), ("16384", 1), ("16384", 0), ("32768", 1)):
        monkeypatch.setenv("ST_EMU_WINDOW", window)
        same_bits(emu.solve(mat, 16400, kernel=2, threads=128, ctas=5, dynamic=dyn, max_iter=3), want)
    monkeypatch.setenv("ST_EMU_WINDOW", "8192")
    same_bits(emu.solve(mat, 16400, kernel=2, threads=64, ctas=3, world=2, dynamic=1, max_iter=2), expect(mat, max_iter=2))
    same_bits(emu.solve(mat, 16400, kernel=2, threads=64, ctas=4, dynamic=1, sweep=0, max_iter=2), expect(mat, max_iter=2))
