"""The relative stop test (ST_STOP_RELATIVE / oracle.STOP_RELATIVE) on the CPU
oracle.  It is an EXTENSION: the reference only has the absolute test
(similarity_transform.cpp:413-421), parity is defined on that one, and the default never changes.
What is pinned here is the extension's own contract, which the CUDA kernels are then held to bit
for bit (tests/test_zz_gpu_bitexact.py):

    converged  <=>  for every r:  |s[r] - s[(r+1) % N]|  <  eps * max(0, max_r s[r])      (strict <)
"""
import numpy as np
import pytest

import oracle

A3 = np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)     # reference tests/test.cpp:84-94


def test_predicate_definition():
    v = np.array([10.0, 10.004, 10.008, 10.0], dtype=np.float32)
    # max = 10.008, eps * max ~ 0.010008: every adjacent difference (0.004, 0.004, 0.008, 0) is below
    assert oracle.stop_relative(v, 1e-3) == 1
    assert oracle.stop(v, 1e-3) == 0                       # the absolute test fails on the same data
    # strict <: a difference equal to the threshold fails
    w = np.array([1024.0, 1025.0], dtype=np.float32)       # max 1025, diff 1 both ways (wrap pair)
    eps = np.float32(1.0) / np.float32(1025.0)
    thr = np.float32(eps) * np.float32(1025.0)
    assert oracle.stop_relative(w, float(eps)) == (1 if np.float32(1.0) < thr else 0)
    assert oracle.stop_relative(w, 2e-3) == 1 and oracle.stop_relative(w, 5e-4) == 0
    # the wrap pair counts
    ramp = (np.float32(100.0) + np.arange(1000, dtype=np.float32) * np.float32(1e-3)).astype(np.float32)
    assert oracle.stop_relative(ramp, 5e-3) == 0           # 100.999 -> 100.0 across the wrap is ~1e-2 relative
    assert oracle.stop_relative(ramp, 2e-2) == 1


def test_nan_never_converges_and_nonpositive_max_never_converges():
    v = np.array([1.0, np.nan, 1.0], dtype=np.float32)
    assert oracle.stop_relative(v, 1.0) == 0
    # max cell starts from 0 (reference :169): all-negative row sums give threshold 0 -> never true
    assert oracle.stop_relative(-np.ones(4, dtype=np.float32), 1e-3) == 0
    mat = (oracle.uniform(16, 3) + np.float32(0.5)).astype(np.float32)
    mat[5, 7] = np.nan
    *_, it = oracle.similarity_transform(mat, max_itr=50, stop=oracle.STOP_RELATIVE)
    assert it == 50


@pytest.mark.parametrize("form", [oracle.FORM_INPLACE, oracle.FORM_READONLY])
def test_scale_invariance(form):
    """Scaling A by a power of two scales every row sum exactly, so the relative test stops in the same
    round with the same (raw) eigenvector and lambda scaled exactly; the absolute test does not."""
    H = oracle.hilbert(512)
    base = oracle.similarity_transform(H, form=form, stop=oracle.STOP_RELATIVE)
    for scale in (2.0 ** -6, 2.0 ** 9):
        got = oracle.similarity_transform((H * np.float32(scale)).astype(np.float32), form=form, stop=oracle.STOP_RELATIVE)
        assert got[3] == base[3]
        assert float(got[0]) == float(base[0]) * scale
        assert np.array_equal(got[1], base[1])
    abs_small = oracle.similarity_transform((H * np.float32(2.0 ** -6)).astype(np.float32), form=form)
    abs_big = oracle.similarity_transform((H * np.float32(2.0 ** 9)).astype(np.float32), form=form)
    assert abs_small[3] < abs_big[3]                       # the reference's test depends on the scale


def test_stops_where_the_absolute_test_cannot():
    """uniform (0,1] 16384^2: one ulp of lambda ~ 8192 is 9.8e-4, the adjacent differences sit at a
    noise floor of ~2.4e-3 > EPS and the reference's test never holds.  The relative test
    with a threshold above the fp32 noise floor (2.4e-3 / 8192 = 3e-7) stops after a few rounds with a
    better eigenpair than the capped run needs."""
    dim = 16384
    U = oracle.uniform(dim, 0x5EED0001)
    *_, it_abs = oracle.similarity_transform(U, form=oracle.FORM_READONLY, max_itr=12)
    assert it_abs == 12                                    # never converges, hits the cap
    val, vec, _, it = oracle.similarity_transform(U, form=oracle.FORM_READONLY, eps=1e-6, stop=oracle.STOP_RELATIVE)
    assert 2 <= it <= 8
    rows = [0, 1, 8191, dim - 1]
    lhs = U[rows].astype(np.float64) @ vec.astype(np.float64)
    assert np.max(np.abs(lhs - float(val) * vec[rows]) / np.abs(lhs)) < 1e-5


def test_known_answers_still_hold_with_a_tight_relative_threshold():
    # 3x3 golden (reference tests/test.cpp:96-102): lambda 7.53114; eps_rel = 1e-3 / 7.5 is the same test there
    val, vec, _, it = oracle.similarity_transform(A3, eps=1e-3 / 7.53114, stop=oracle.STOP_RELATIVE)
    assert abs(val - 7.53114) < 1e-3 and it in (4, 5)
    assert np.allclose(vec / vec.max(), np.array([0.394074, 0.578844, 0.997451]) / 0.997451, atol=1e-3)


def test_every_summation_order_and_rank_count_agree_on_the_round():
    H = oracle.hilbert(1024)
    base = oracle.similarity_transform(H, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, stop=oracle.STOP_RELATIVE)
    for sm in (oracle.SUM_SEQUENTIAL, oracle.SUM_LANES16, oracle.SUM_SUBGROUP32):
        got = oracle.similarity_transform(H, form=oracle.FORM_READONLY, sum_mode=sm, stop=oracle.STOP_RELATIVE)
        assert got[3] == base[3] and abs(float(got[0]) - float(base[0])) <= 1e-5 * float(base[0])
    for ranks in (2, 8):
        got = oracle.similarity_transform(H, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, stop=oracle.STOP_RELATIVE, ranks=ranks)
        assert got[3] == base[3] and got[0] == base[0] and np.array_equal(got[1], base[1])
