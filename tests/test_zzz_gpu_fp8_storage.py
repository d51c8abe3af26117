"""fp8 STORAGE of the matrix (opt-in extension) on the GPU.

The matrix is held as one byte per element (e4m3) plus ONE power-of-two fp32 scale per row -- a quarter of the HBM
bytes per round -- while the eigenvector, the row sums and every accumulation stay fp32.  It changes results, so it
is outside reference parity; its own contract is exact all the same: e4m3 -> fp32 is exact and the scales are powers
of two, so an fp8-storage solve must return the bits of an fp32 solve of the dequantised matrix in the fp32 kernels'
own order (the kernels reduce 4-element words), which is what the oracle computes with to_fp8_rows() + SUM_CUDA (pinned on the CPU in
tests/test_oracle_cuda_order.py; the kernels' logic is bit-identical to it on the emulation harness,
tests/test_kernel_logic_emulated.py).  The quantiser is checked code for code: the device's
cvt.rn.satfinite.e4m3x2.f32 against the oracle's nearest-code search.
"""
import numpy as np
import pytest

import oracle
from eigen_value_b200 import FORM_INPLACE, STOP_RELATIVE

pytestmark = pytest.mark.gpu


def _same_bits(info, vec, want):
    w_val, w_vec, _, w_it = want
    assert info.iter_count == w_it, (info.iter_count, w_it)
    assert np.float32(info.eigen_val).view(np.uint32) == np.float32(w_val).view(np.uint32), (float(info.eigen_val), float(w_val))
    assert np.array_equal(vec.view(np.uint32), w_vec.view(np.uint32))


def _case(kind, dim):
    mat = oracle.hilbert(dim) if kind == "hilbert" else (oracle.uniform(dim, 3000 + dim) + np.float32(0.25)).astype(np.float32)
    back, codes, scale = oracle.to_fp8_rows(mat)
    return mat, back, codes, scale


def _want(back, **kw):
    return oracle.similarity_transform(back, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_FP8, **kw)


def test_device_conversion_picks_the_same_codes_and_scales(solver):
    rng = np.random.default_rng(8)
    x = (rng.random((300, 256)) * np.exp(rng.normal(0, 6, (300, 1)))).astype(np.float32)
    x[3] = 0                                                  # an all-zero row keeps scale 1
    x[5, :4] = [448.0, -448.0, 1e-9, -3.0]                    # saturation edge, underflow to zero, sign
    x[7, 9] = np.nan
    x[11] *= np.float32(1e30)
    x[12] *= np.float32(1e-30)
    vals = oracle.fp8_e4m3_values()                           # every code and every midpoint between two codes (ties)
    mids = ((vals[:-1].astype(np.float64) + vals[1:]) / 2).astype(np.float32)
    x[20:22] = 0
    x[20, :127] = vals
    x[20, 127] = 448.0
    x[21, :126] = mids
    x[21, 126] = 448.0
    back, codes, scale = oracle.to_fp8_rows(x)
    d_codes, d_scale = solver.to_fp8(solver.upload(x), 300, 256)
    solver.synchronize()
    assert np.array_equal(d_scale.download(np.float32, 300), scale, equal_nan=True) and np.isnan(scale[7])   # the NaN row
    assert np.array_equal(d_codes.download(np.uint8, 300 * 256).reshape(300, 256), codes)


@pytest.mark.parametrize("kind", ["hilbert", "uniform"])
@pytest.mark.parametrize("dim", [4, 16, 64, 508, 1008, 1024, 4096, 8192, 8196, 16384])
def test_fp8_storage_solve_is_bit_identical_to_the_oracle(solver, kind, dim):
    # resident-e kernel (configuration 11: no prefetch slots); 8196 / 16384: rows of two work units
    mat, back, codes, scale = _case(kind, dim)
    d8, dsc = solver.upload(codes), solver.upload(scale)
    cap = 12 if dim > 4100 else 1000
    info, vec = solver.solve_device(d8, dim, fp8_scale=dsc, max_iter=cap)
    assert info.kernel_id == 11 and info.bytes_per_round == dim * dim
    _same_bits(info, vec, _want(back, max_itr=cap))
    # and near the fp32-storage answer: every entry moved by at most 2^-4 relative (or 2^-10 of its row's scale)
    full, _ = solver.solve_device(solver.upload(mat), dim, max_iter=cap)
    assert abs(float(info.eigen_val) - float(full.eigen_val)) <= 0.04 * float(full.eigen_val)


def test_device_converted_storage_equals_host_converted(solver):
    dim = 2048
    mat, back, codes, scale = _case("hilbert", dim)
    d8, dsc = solver.to_fp8(solver.hilbert(dim), dim, dim)
    solver.synchronize()
    assert np.array_equal(d8.download(np.uint8, dim * dim).reshape(dim, dim), codes)
    assert np.array_equal(dsc.download(np.float32, dim), scale)
    info, vec = solver.solve_device(d8, dim, fp8_scale=dsc)
    _same_bits(info, vec, _want(back))


@pytest.mark.parametrize("dim", [1008, 8192, 9216])
def test_general_loop_on_fp8_storage(solver, dim):
    mat, back, codes, scale = _case("uniform", dim)
    info, vec = solver.solve_device(solver.upload(codes), dim, fp8_scale=solver.upload(scale), kernel=1, max_iter=6)
    assert info.kernel_id == 1
    _same_bits(info, vec, _want(back, max_itr=6))


def test_general_loop_beyond_the_resident_limit(solver):
    dim = 40960                                            # 1.6 GiB of fp8; the oracle holds the dequantised fp32 copy
    if solver.hbm_bytes < 16 * 2**30:
        pytest.skip("needs ~10 GiB of device memory")
    d32 = solver.uniform(dim, 0x5EED0003)
    d8, dsc = solver.to_fp8(d32, dim, dim)
    solver.synchronize()
    d32.free()
    info, vec = solver.solve_device(d8, dim, fp8_scale=dsc, max_iter=3)
    d8.free()
    assert info.kernel_id == 1
    back = oracle.to_fp8_rows(oracle.uniform(dim, 0x5EED0003))[0]
    _same_bits(info, vec, _want(back, max_itr=3))


def test_relative_stop_on_fp8_storage(solver):
    dim = 4096
    mat, back, codes, scale = _case("uniform", dim)
    info, vec = solver.solve_device(solver.upload(codes), dim, fp8_scale=solver.upload(scale), eps=1e-6, stop=STOP_RELATIVE)
    _same_bits(info, vec, _want(back, eps=1e-6, stop=oracle.STOP_RELATIVE))


def test_a_nan_in_the_matrix_travels_in_its_rows_scale_and_runs_to_the_cap(solver):
    # the kernels read a code's magnitude bits as a number, so the codes cannot carry a NaN: the row's scale does, and
    # that row's sum is NaN in every round -- the stop test can never hold (like the fp32 path, tests/test_gpu_parity.py)
    dim = 64
    mat = oracle.hilbert(dim).copy()
    mat[5, 7] = np.nan
    d8, dsc = solver.to_fp8(solver.upload(mat), dim, dim)
    info, vec = solver.solve_device(d8, dim, fp8_scale=dsc, max_iter=7)
    assert info.iter_count == 7 and info.passes == 7
    assert np.isnan(dsc.download(np.float32, dim)[5]) and np.isnan(vec).any()


def test_unsupported_combinations_are_refused(solver):
    _, _, codes, scale = _case("hilbert", 64)
    d, dsc = solver.upload(codes), solver.upload(scale)
    with pytest.raises(Exception):
        solver.solve_device(d, 62, fp8_scale=dsc)                       # dim % 4 != 0
    with pytest.raises(Exception):
        solver.solve_device(d, 64, fp8_scale=dsc, form=FORM_INPLACE)    # read-only form only
    with pytest.raises(Exception):
        solver.solve_device(d, 64, fp8_scale=dsc, kernel=13)            # prefetching configuration is fp32-only
    with pytest.raises(Exception):
        solver.solve_device(d, 64, fp8_scale=dsc, accumulate=1)         # fp32 accumulation only
    info, _ = solver.solve_device(d, 64, fp8_scale=dsc)                 # the handle stays usable
    assert info.iter_count >= 1
