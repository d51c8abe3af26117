"""bf16 STORAGE of the matrix (opt-in extension) on the GPU.

The matrix is held as bfloat16 -- half the HBM bytes per round -- while the eigenvector, the row
sums and every accumulation stay fp32.  It changes results, so it is outside reference parity; its own
contract is exact all the same: bf16 -> fp32 is exact, so a bf16-storage solve must return the bits of
an fp32 solve of the bf16-rounded matrix in the fp32 kernels' own order (the kernels reduce 4-element words), which
is what the oracle computes with to_bf16() + SUM_CUDA (pinned on the CPU in tests/test_oracle_cuda_order.py).

Its logic also runs on the CPU emulation harness (tests/test_kernel_logic_emulated.py: both kernels, several work
units per row, sharded, relative stop, the conversion kernel), bit-identical to the oracle there.
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
    mat = oracle.hilbert(dim) if kind == "hilbert" else (oracle.uniform(dim, 2000 + dim) + np.float32(0.25)).astype(np.float32)
    rounded, bits = oracle.to_bf16(mat)
    return mat, rounded, bits


def test_device_conversion_is_round_to_nearest_even(solver):
    for n in (1, 7, 8, 1000, 4096 * 33 + 5):
        x = (np.random.default_rng(n).random(n) * 100 + 1e-3).astype(np.float32)
        x[: min(n, 4)] = np.array([1.00390625, 1.01171875, 1.0, 3.0e38], dtype=np.float32)[: min(n, 4)]   # ties, large
        d = solver.to_bf16(solver.upload(x), n)
        solver.synchronize()
        assert np.array_equal(d.download(np.uint16, n), oracle.to_bf16(x)[1])


@pytest.mark.parametrize("kind", ["hilbert", "uniform"])
@pytest.mark.parametrize("dim", [4, 64, 508, 1000, 1024, 4096, 8192, 8196, 16384])
def test_bf16_storage_solve_is_bit_identical_to_the_oracle(solver, kind, dim):
    # resident-e kernel (configuration 11: no prefetch slots); 8196 / 16384: rows of two work units
    mat, rounded, bits = _case(kind, dim)
    d16 = solver.upload(bits)
    # large uniform matrices never meet the reference's absolute stop test in fp32: cap the rounds
    cap = 12 if (kind == "uniform" and dim > 4100) else 1000
    info, vec = solver.solve_device(d16, dim, bf16=True, max_iter=cap)
    assert info.kernel_id == 11 and info.bytes_per_round == 2 * dim * dim
    _same_bits(info, vec, oracle.similarity_transform(rounded, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_BF16,
                                                      max_itr=cap))
    # and close to the fp32-storage answer: the entries moved by <= 2^-9 relative
    full, _ = solver.solve_device(solver.upload(mat), dim, max_iter=cap)
    assert abs(float(info.eigen_val) - float(full.eigen_val)) <= 2.0 ** -8 * float(full.eigen_val)


def test_device_converted_storage_equals_host_converted(solver):
    dim = 2048
    mat, rounded, bits = _case("hilbert", dim)
    d16 = solver.to_bf16(solver.hilbert(dim), dim * dim)
    solver.synchronize()
    assert np.array_equal(d16.download(np.uint16, dim * dim).reshape(dim, dim), bits)
    info, vec = solver.solve_device(d16, dim, bf16=True)
    _same_bits(info, vec, oracle.similarity_transform(rounded, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_BF16))


@pytest.mark.parametrize("dim", [1000, 8192, 9216])
def test_general_loop_on_bf16_storage(solver, dim):
    mat, rounded, bits = _case("uniform", dim)
    info, vec = solver.solve_device(solver.upload(bits), dim, bf16=True, kernel=1, max_iter=6)
    assert info.kernel_id == 1
    _same_bits(info, vec, oracle.similarity_transform(rounded, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_BF16, max_itr=6))


def test_general_loop_beyond_the_resident_limit(solver):
    dim = 40960                                            # 3.1 GiB of bf16; the oracle holds the rounded fp32 copy
    if solver.hbm_bytes < 16 * 2**30:
        pytest.skip("needs ~10 GiB of device memory")
    d32 = solver.uniform(dim, 0x5EED0003)
    d16 = solver.to_bf16(d32, dim * dim)
    solver.synchronize()
    d32.free()
    info, vec = solver.solve_device(d16, dim, bf16=True, max_iter=3)
    d16.free()
    assert info.kernel_id == 1
    rounded = oracle.to_bf16(oracle.uniform(dim, 0x5EED0003))[0]
    _same_bits(info, vec, oracle.similarity_transform(rounded, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_BF16, max_itr=3))


def test_relative_stop_on_bf16_storage(solver):
    dim = 4096
    mat, rounded, bits = _case("uniform", dim)
    info, vec = solver.solve_device(solver.upload(bits), dim, bf16=True, eps=1e-6, stop=STOP_RELATIVE)
    _same_bits(info, vec, oracle.similarity_transform(rounded, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_BF16,
                                                      eps=1e-6, stop=oracle.STOP_RELATIVE))


def test_unsupported_combinations_are_refused(solver):
    bits = oracle.to_bf16(oracle.hilbert(64))[1]
    d = solver.upload(bits)
    with pytest.raises(Exception):
        solver.solve_device(d, 62, bf16=True)                       # dim % 4 != 0
    with pytest.raises(Exception):
        solver.solve_device(d, 64, bf16=True, form=FORM_INPLACE)    # read-only form only
    with pytest.raises(Exception):
        solver.solve_device(d, 64, bf16=True, kernel=13)            # prefetching configuration is fp32-only
    info, _ = solver.solve_device(d, 64, bf16=True)                 # the handle stays usable
    assert info.iter_count >= 1
