"""GPU parity, strongest form: the CUDA path against the CPU oracle BIT FOR BIT.

tests/test_gpu_parity.py holds the CUDA path to BASELINE.json's tolerances against the oracle in
the reference's summation orders.  Every CUDA round kernel evaluates a row in one fixed order
(DESIGN.md section 2); oracle.SUM_CUDA restates that order on the CPU (oracle.c:
row_dot_cuda_order, pinned in tests/test_oracle_cuda_order.py -- among others against eigenvalues
recorded on B200s in round 1), so here eigenvalue, raw eigenvector and round count must be
IDENTICAL, for every kernel the automatic choice can pick, both forms, vector and scalar loads.

(The file name sorts last on purpose: these are the newest assertions of the suite and must not
shadow the tolerance-parity tests under `pytest -x`.)
"""
import numpy as np
import pytest

import oracle
from eigen_value_b200 import EigenValue, FORM_INPLACE, FORM_READONLY

pytestmark = pytest.mark.gpu

A3 = np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)     # reference tests/test.cpp:84-94


def _oracle(mat, form=FORM_READONLY, **kw):
    o_form = oracle.FORM_READONLY if form == FORM_READONLY else oracle.FORM_INPLACE
    val, vec, _, it = oracle.similarity_transform(mat, form=o_form, sum_mode=oracle.SUM_CUDA, **kw)
    return val, vec, it


def _assert_same_bits(got, want, what=""):
    g_val, g_vec, g_it = got
    w_val, w_vec, w_it = want
    assert g_it == w_it, (what, g_it, w_it)
    assert np.float32(g_val).view(np.uint32) == np.float32(w_val).view(np.uint32), (what, float(g_val), float(w_val))
    assert np.array_equal(g_vec.view(np.uint32), w_vec.view(np.uint32)), what


def _matrix(kind, dim):
    if kind == "hilbert":
        return oracle.hilbert(dim)
    return (oracle.uniform(dim, seed=1000 + dim) + np.float32(0.25)).astype(np.float32)


@pytest.fixture(scope="module")
def ev():
    return EigenValue()


# on-chip cluster kernel (dim % 4 == 0, <= 512), resident-e kernel (dim <= 32768; dim % 4 != 0 on its scalar-unit
# build, configuration 11); 8196 / 12288: rows of two work units
DIMS = [1, 2, 3, 4, 5, 31, 33, 100, 128, 257, 512, 640, 1000, 1023, 1024, 2048, 4100, 8192, 8196, 12288]


@pytest.mark.parametrize("kind", ["hilbert", "uniform"])
@pytest.mark.parametrize("dim", DIMS)
def test_drop_in_boundary_is_bit_identical_to_the_oracle(ev, kind, dim):
    if kind == "uniform" and dim > 4100:
        # uniform matrices of this size never satisfy the reference's absolute stop test in fp32:
        # 1000 rounds cost the CPU oracle minutes; they go through st_solve_device with a round cap below
        pytest.skip("covered with a round cap by test_large_uniform_matrices_with_a_round_cap")
    mat = _matrix(kind, dim)
    val, vec, ms, it = ev.similarity_transform(mat)                     # make_queue + max_eigen_value
    _assert_same_bits((val, vec, it), _oracle(mat), f"{kind}-{dim}")


@pytest.mark.parametrize("dim", [8192, 8196, 12288])
def test_large_uniform_matrices_with_a_round_cap(solver, dim):
    mat = _matrix("uniform", dim)
    info, vec = solver.solve_device(solver.upload(mat), dim, max_iter=12)
    assert info.iter_count == 12                                          # the absolute test cannot hold here
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, max_itr=12), f"uniform-{dim} capped")


def test_three_by_three_golden_bit_identical(ev):
    val, vec, ms, it = ev.similarity_transform(A3)
    assert it == 4
    _assert_same_bits((val, vec, it), _oracle(A3), "3x3 golden")


@pytest.mark.parametrize("dim", [3, 100, 1000, 2048, 4100])
def test_in_place_form_is_bit_identical_to_the_oracle(solver, dim):
    # the literal W <- D^-1 W D rescale (reference similarity_transform.cpp:324-325): same ops, same order
    mat = _matrix("hilbert", dim) if dim > 3 else A3
    info, vec = solver.solve_device(solver.upload(mat), dim, form=FORM_INPLACE)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, FORM_INPLACE), f"in-place {dim}")


@pytest.mark.parametrize("kernel", [1, 13, 11])
def test_every_kernel_family_is_bit_identical_to_the_oracle(solver, kernel):
    # general loop, resident-e with and without the cross-barrier prefetch, at a size where
    # rows span two work units (two 8192-column chunks)
    dim = 9216
    mat = _matrix("uniform", dim)
    info, vec = solver.solve_device(solver.upload(mat), dim, kernel=kernel, max_iter=8)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, max_itr=8), f"kernel {kernel}")


@pytest.mark.parametrize("dim", [5, 1001, 4099, 8191, 8195])
def test_ragged_dims_run_on_the_resident_e_kernel_and_match_the_general_loop(solver, dim):
    # dim % 4 != 0: automatic = resident-e configuration 11 on scalar units (dynamic work units, e in shared memory);
    # the general loop (kernel 1) is the other kernel built for these dimensions -- same order, same bits, = the oracle
    mat = _matrix("uniform", dim)
    cap = 6 if dim > 4100 else 1000
    d = solver.upload(mat)
    info, vec = solver.solve_device(d, dim, max_iter=cap)
    base, base_vec = solver.solve_device(d, dim, max_iter=cap, kernel=1)
    assert info.kernel_id == 11 and base.kernel_id == 1
    want = _oracle(mat, max_itr=cap)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), want, f"ragged {dim} resident-e")
    _assert_same_bits((base.eigen_val, base_vec, base.iter_count), want, f"ragged {dim} general")


def test_multi_chunk_16384_bit_identical(solver):
    dim = 16384
    info, vec = solver.solve_device(solver.hilbert(dim), dim)
    assert info.iter_count == 18
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(oracle.hilbert(dim)), "hilbert-16384")


def test_wide_and_general_loops_beyond_the_resident_limit_bit_identical(solver):
    # N > 32768: the wide kernel (automatic; two windows of 32768 + 8192 columns, 5 chunks per row) and the general
    # column-chunked loop (kernel 1); capped at 3 rounds so the oracle needs ~10 s for the 6.25 GiB matrix
    dim = 40960
    if solver.hbm_bytes < 16 * 2**30:
        pytest.skip("needs 6.25 GiB of device memory")
    d = solver.uniform(dim, 0x5EED0003)
    info, vec = solver.solve_device(d, dim, max_iter=3)
    base, base_vec = solver.solve_device(d, dim, max_iter=3, kernel=1)
    d.free()
    assert info.kernel_id == 2 and base.kernel_id == 1
    want = oracle.similarity_transform_generated("uniform", dim, 0x5EED0003, max_itr=3)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), (want[0], want[1], want[3]), "uniform-40960 wide")
    _assert_same_bits((base.eigen_val, base_vec, base.iter_count), (want[0], want[1], want[3]), "uniform-40960 general")


@pytest.mark.parametrize("dim", [4, 520, 2048, 8200, 16384])
def test_wide_kernel_at_sizes_the_resident_kernel_also_takes(solver, dim):
    # explicit kernel 2 below the resident limit: one window, rows of one or two chunks -- the bits of every other kernel
    mat = _matrix("uniform", dim)
    info, vec = solver.solve_device(solver.upload(mat), dim, kernel=2, max_iter=8)
    assert info.kernel_id == 2
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, max_itr=8), f"wide {dim}")


@pytest.mark.parametrize("dim", [7, 64, 1000, 2048, 8200])
def test_standalone_row_sum_kernel_bit_identical(solver, dim):
    # st_sum_across_rows (reference similarity_transform.cpp:77-152) shares the fused kernels' order
    m = oracle.uniform(dim, seed=99 + dim)
    assert np.array_equal(solver.sum_across_rows(m), oracle.sum_across_rows(m, oracle.SUM_CUDA))


def test_eps_and_cap_options_bit_identical(solver):
    dim = 1024
    mat = oracle.hilbert(dim)
    d = solver.upload(mat)
    info, vec = solver.solve_device(d, dim, eps=1e-2)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, eps=1e-2), "eps 1e-2")
    info, vec = solver.solve_device(d, dim, max_iter=5)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), _oracle(mat, max_itr=5), "cap 5")


# ---- ST_STOP_RELATIVE (extension): same bits as the oracle's relative test ----------
from eigen_value_b200 import STOP_RELATIVE  # noqa: E402


@pytest.mark.parametrize("form", [FORM_READONLY, FORM_INPLACE])
@pytest.mark.parametrize("dim", [3, 100, 512, 1000, 1023, 2048, 8192])
def test_relative_stop_is_bit_identical_to_the_oracle(solver, form, dim):
    # cluster kernel (<= 512), resident-e kernel, general loop with vector and scalar loads, both forms
    mat = _matrix("uniform", dim) if dim > 3 else A3
    for eps in (1e-3, 1e-6):
        info, vec = solver.solve_device(solver.upload(mat), dim, form=form, eps=eps, stop=STOP_RELATIVE, max_iter=60)
        want = _oracle(mat, form, eps=eps, stop=oracle.STOP_RELATIVE, max_itr=60)
        _assert_same_bits((info.eigen_val, vec, info.iter_count), want, f"relative {dim} eps={eps}")


def test_relative_stop_converges_where_the_reference_test_cannot(solver):
    """uniform (0,1] 16384^2: the reference's absolute test never holds in fp32 and the loop
    runs to the cap; the relative test stops after a few rounds with A.v ~= lambda.v to 1e-5."""
    dim = 16384
    d = solver.uniform(dim, 0x5EED0001)
    capped, _ = solver.solve_device(d, dim, max_iter=12)
    assert capped.iter_count == 12
    info, vec = solver.solve_device(d, dim, eps=1e-6, stop=STOP_RELATIVE)
    mat = oracle.uniform(dim, 0x5EED0001)
    _assert_same_bits((info.eigen_val, vec, info.iter_count),
                      _oracle(mat, eps=1e-6, stop=oracle.STOP_RELATIVE), "uniform-16384 relative")
    assert 2 <= info.iter_count <= 8
    rows = [0, 1, 8191, dim - 1]
    lhs = mat[rows].astype(np.float64) @ vec.astype(np.float64)
    assert np.max(np.abs(lhs - float(info.eigen_val) * vec[rows]) / np.abs(lhs)) < 1e-5


def test_relative_stop_with_nan_hits_the_cap_and_unknown_options_are_refused(solver):
    mat = (oracle.uniform(64, 5) + np.float32(0.5)).astype(np.float32)
    mat[32, 21] = np.nan
    info, _ = solver.solve_device(solver.upload(mat), 64, stop=STOP_RELATIVE, max_iter=50)
    assert info.iter_count == 50
    d = solver.hilbert(1024)
    with pytest.raises(Exception):
        solver.solve_device(d, 1024, kernel=5, stop=STOP_RELATIVE)     # a kernel id that does not exist (round 1: TMA ring)
    with pytest.raises(Exception):
        solver.solve_device(d, 1024, stop=7)                           # unknown mode
    info, _ = solver.solve_device(d, 1024)                             # the handle stays usable
    assert info.iter_count == 13


# ---- ST_ACC_F64 (extension): fp64 accumulators, same order; same bits as the oracle's SUM_CUDA_F64 -----------
from eigen_value_b200 import ACC_F64  # noqa: E402


@pytest.mark.parametrize("dim", [3, 100, 512, 1000, 1023, 2048, 8192, 8200])
def test_fp64_accumulation_is_bit_identical_to_the_oracle(solver, dim):
    # resident-e kernel (N % 4 == 0; the on-chip cluster kernel is not built for it), general loop otherwise
    mat = _matrix("uniform", dim) if dim > 3 else A3
    d = solver.upload(mat)
    cap = 12 if dim > 4100 else 1000                     # large uniform matrices never meet the absolute test
    info, vec = solver.solve_device(d, dim, accumulate=ACC_F64, max_iter=cap)
    want = oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_F64, max_itr=cap)
    _assert_same_bits((info.eigen_val, vec, info.iter_count), (want[0], want[1], want[3]), f"fp64 accumulation {dim}")
    base, _ = solver.solve_device(d, dim, max_iter=cap)
    assert abs(float(info.eigen_val) - float(base.eigen_val)) <= 1e-5 * float(base.eigen_val)
    if dim % 4 == 0 and dim <= 2048:
        info, vec = solver.solve_device(d, dim, accumulate=ACC_F64, kernel=1, eps=1e-6, stop=STOP_RELATIVE, max_iter=60)
        want = oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA_F64, eps=1e-6,
                                           stop=oracle.STOP_RELATIVE, max_itr=60)
        _assert_same_bits((info.eigen_val, vec, info.iter_count), (want[0], want[1], want[3]), f"fp64 + relative {dim}")


def test_fp64_accumulation_refusals(solver):
    d = solver.hilbert(256)
    for bad in (dict(form=FORM_INPLACE), dict(kernel=2), dict(kernel=20), dict(kernel=11), dict(accumulate=5)):   # 2: wide kernel, fp32 accumulators only
        kw = dict(accumulate=ACC_F64)
        kw.update(bad)
        with pytest.raises(Exception):
            solver.solve_device(d, 256, **kw)
    info, _ = solver.solve_device(d, 256, accumulate=ACC_F64)
    assert info.iter_count == 10 and info.kernel_id == 13          # README.md:71; resident-e instead of the cluster kernel


def test_eigenvalues_printed_by_the_b200_in_round_1_are_reproduced(ev):
    """profiles/r1_reference_format_table.txt was produced on a B200 in round 1 (tools/bench_table.py).  Seven decimals
    identify one float32 in [2, 4): whatever runs this test -- the B200 again, or the emulated library on the CPU --
    must print the same eigenvalues and round counts."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r1_reference_format_table.txt")
    rows = re.findall(r"^(\d+)\s*x\s*\d+\s+[\d.]+ ms\s+(\d+) round\(s\)\s+max_eigen_value\(\):.*lambda = ([\d.]+)", open(path).read(), re.M)
    assert len(rows) == 7
    for n, rounds, lam in rows:
        n = int(n)
        if n > 2048:
            continue                                          # kept cheap: the larger sizes are covered above
        val, vec, ms, it = ev.similarity_transform(oracle.hilbert(n))
        assert it == int(rounds) and "%.7f" % float(val) == lam, (n, it, float(val), lam)
