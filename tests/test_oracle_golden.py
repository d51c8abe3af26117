"""Pins the CPU oracle (oracle/oracle.c) to every known answer the reference holds for the
similarity_transform() path.  CPU only; runs in the build container and on the GPU box alike.

Sources (paths relative to the reference repository):
  tests/test.cpp:22-104   per-kernel fixtures + the 3x3 golden eigenpair
  utils.cpp:5-122         the fixture generators
  main.py:52-58           the same 3x3 golden for the sequential model
  README.md:70-76         Hilbert round counts (identical on all six published devices)
  wrapper/python/test.py  A.v ~= lambda.v acceptance criterion
"""
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

A3 = np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)     # tests/test.cpp:84-94
HILBERT_ROUNDS = {128: 9, 256: 10, 512: 12, 1024: 13, 2048: 14, 4096: 15, 8192: 17}  # README.md:70-76

FORMS = [oracle.FORM_INPLACE, oracle.FORM_READONLY]
SUMS = [oracle.SUM_SEQUENTIAL, oracle.SUM_LANES16, oracle.SUM_SUBGROUP32]


# ---- per-kernel fixtures, tests/test.cpp:22-73 with N = 1024 (:7) ---------------------------
N = 1 << 10


@pytest.mark.parametrize("sum_mode", SUMS)
def test_sum_across_rows_identity(sum_mode):
    # tests/test.cpp:22-30, checker utils.cpp:29-35: every row sum of I is exactly 1
    s = oracle.sum_across_rows(np.eye(N, dtype=np.float32), sum_mode)
    assert np.all(s == np.float32(1.0))


def test_find_max_of_iota():
    # tests/test.cpp:32-41, generator utils.cpp:37-59: v[r] = r + 1 -> max == N
    v = np.arange(1, N + 1, dtype=np.float32)
    assert oracle.find_max(v) == float(N)


def test_find_max_starts_from_zero():
    # similarity_transform.cpp:169 fills the max cell with 0 first: all-negative input gives 0
    assert oracle.find_max(-np.ones(8, dtype=np.float32)) == 0.0


def test_compute_eigen_vector_first_update():
    # tests/test.cpp:43-54, checker utils.cpp:61-72: after init + one update e[r] == v[r]/max
    v = np.arange(1, N + 1, dtype=np.float32)
    e = np.empty(N, dtype=np.float32)
    oracle.lib().oracle_initialise_eigen_vector(e, N)
    assert np.all(e == 1.0)
    oracle.compute_eigen_vector(v, float(N), e)
    assert np.max(np.abs(v / np.float32(N) - e)) == 0.0


def test_stop_success_data():
    # tests/test.cpp:56-64, generator utils.cpp:74-97: all entries 1 + 1e-4 -> converged
    v = np.full(N, np.float32(1.0) + np.float32(1e-4), dtype=np.float32)
    assert oracle.stop(v) == 1


def test_stop_fail_data_fails_only_through_wrap_pair():
    # tests/test.cpp:66-73, generator utils.cpp:99-122: v[r] = (r+1)*1e-4.  All interior
    # differences are 1e-4 < EPS; only |v[N-1] - v[0]| = 0.1023 breaks it (circular test,
    # similarity_transform.cpp:413-421).
    v = (np.arange(1, N + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    assert np.all(np.abs(np.diff(v)) < oracle.EPS)
    assert oracle.stop(v) == 0


def test_stop_is_strict_less_than():
    v = np.array([0.0, 1e-3, 0.0, 0.0], dtype=np.float32)
    assert oracle.stop(v, eps=float(np.float32(1e-3))) == 0
    assert oracle.stop(v, eps=float(np.nextafter(np.float32(1e-3), np.float32(1)))) == 1


def test_compute_next_matrix_order_of_operations():
    # similarity_transform.cpp:324-325: W[r][c] *= (1.f / s[r]) * s[c]
    rng = np.random.default_rng(7)
    W = rng.random((16, 16), dtype=np.float32) + np.float32(0.5)
    s = oracle.sum_across_rows(W)
    expect = W * ((np.float32(1.0) / s)[:, None] * s[None, :])
    oracle.compute_next_matrix(W, s)
    assert np.array_equal(W, expect.astype(np.float32))


def test_dim_one_converges_at_round_zero():
    val, vec, _, it = oracle.similarity_transform(np.array([[3.0]], dtype=np.float32))
    assert (float(val), it, float(vec[0])) == (3.0, 0, 1.0)


# ---- the 3x3 golden eigenpair ----------------------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("sum_mode", SUMS)
def test_three_by_three_golden(form, sum_mode):
    # tests/test.cpp:96-102 (and main.py:52-58 to 4 digits): each within EPS = 1e-3
    val, vec, _, it = oracle.similarity_transform(A3, form=form, sum_mode=sum_mode)
    assert abs(val - 7.53114) < 1e-3
    assert abs(vec[0] - 0.394074) < 1e-3
    assert abs(vec[1] - 0.578844) < 1e-3
    assert abs(vec[2] - 0.997451) < 1e-3
    assert it == 4
    # much tighter than the reference asks: the published digits are all significant
    assert abs(val - 7.53114) < 5e-6 and np.max(np.abs(vec - [0.394074, 0.578844, 0.997451])) < 2e-6


# ---- Hilbert round counts ---------------------------------------------------------------------
@pytest.mark.parametrize("dim", sorted(HILBERT_ROUNDS))
def test_hilbert_round_counts_match_published(dim):
    H = oracle.hilbert(dim)
    got = {}
    for form in FORMS:
        val, vec, _, it = oracle.similarity_transform(H, form=form)
        got[form] = (val, vec / vec.max())
        assert it == HILBERT_ROUNDS[dim], (dim, form, it)
    # the two forms are the same iteration up to rounding
    assert abs(got[0][0] - got[1][0]) / got[0][0] < 2e-6
    assert np.max(np.abs(got[0][1] - got[1][1])) < 2e-6


def test_hilbert_generator_formula():
    # utils.cpp:150: 1.f / (float)(r + c + 1)
    H = oracle.hilbert(64)
    r, c = np.indices((64, 64))
    assert np.array_equal(H, (np.float32(1.0) / (r + c + 1).astype(np.float32)).astype(np.float32))
    assert np.array_equal(oracle.hilbert(64, row0=16, rows=8), H[16:24])


@pytest.mark.parametrize("dim", [1024])
@pytest.mark.parametrize("sum_mode", SUMS)
def test_summation_order_does_not_move_the_answer(dim, sum_mode):
    # the reference's own summation order is unspecified (float atomics,
    # similarity_transform.cpp:124-146); every legal order must give the published count
    val, vec, _, it = oracle.similarity_transform(oracle.hilbert(dim), sum_mode=sum_mode)
    assert it == HILBERT_ROUNDS[dim]
    assert abs(val - 2.4455502) / 2.4455502 < 5e-6


# ---- acceptance criterion of wrapper/python/test.py ---------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reference_acceptance_criterion_random_1024(seed):
    # test.py:10-16: random [0,1) float32 1024^2, A v ~= lambda v with atol 1e-3
    rng = np.random.default_rng(seed)
    mat = rng.random((1024, 1024)).astype("f")
    val, vec, _, it = oracle.similarity_transform(mat)
    assert np.all(np.isclose(np.matmul(mat, vec), val * vec, atol=1e-3))
    assert 1 <= it < 10


# ---- sharding must not change a single bit ---------------------------------------------------------
@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("ranks", [2, 3, 8])
def test_row_block_sharding_is_bitwise_neutral(form, ranks):
    mat = oracle.uniform(250, seed=0x5EED0001)
    base = oracle.similarity_transform(mat, form=form)
    shard = oracle.similarity_transform(mat, form=form, ranks=ranks)
    assert base[0] == shard[0] and base[3] == shard[3] and np.array_equal(base[1], shard[1])


# ---- Philox generator -----------------------------------------------------------------------------
def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32 10, counter 0, key 0
    assert [int(x) for x in oracle.philox_block(0, 0)] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_uniform_is_open_closed_and_shard_independent():
    full = oracle.uniform(37, seed=0x5EED0002)
    assert full.min() > 0.0 and full.max() <= 1.0
    assert np.array_equal(oracle.uniform(37, seed=0x5EED0002, row0=5, rows=9), full[5:14])
    w = oracle.philox_block(0, 0x5EED0002)
    assert full[0, 1] == np.float32(((int(w[1]) >> 8) + 1) * 2.0 ** -24)


def test_uniform_large_never_converges_quickly():
    # near the fp32 noise floor the absolute stop test flips on rounding; at
    # N=2048 the loop still ends within a handful of rounds
    val, vec, _, it = oracle.similarity_transform(oracle.uniform(2048, seed=0x5EED0001))
    assert it < 10 and abs(val - 1024) < 20


# ---- fixtures generated from the reference itself (tests/golden/make_golden.py) -----------------------
def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    with open(path) as f:
        return json.load(f)


def test_against_reference_main_py_fixture():
    """main.py (fp64, non-circular stop, reference main.py:25-47) converges to the same
    eigenpair; its own stop rule fires earlier, so only a loose bound is asserted."""
    g = _load("main_py.json")
    for case in g["cases"]:
        mat = np.array(case["matrix"], dtype=np.float32)
        val, vec, _, it = oracle.similarity_transform(mat)
        assert abs(val - case["eigen_val"]) / case["eigen_val"] < 2e-3
        ref = np.array(case["eigen_vec"])
        assert np.max(np.abs(vec / vec.max() - ref / ref.max())) < 2e-3


def test_sequential_model_restatement_reproduces_main_py_bit_for_bit():
    """oracle/sequential.py restates reference main.py:13-47 with the same numpy operations: on the fixtures
    generated from the unmodified main.py it must return the same value, vector and round count exactly
    (same numpy -> same bits; another numpy build may reorder a BLAS sum, then 1e-6 relative)."""
    from oracle import sequential
    g = _load("main_py.json")
    same_numpy = g.get("numpy") == np.__version__
    for case in g["cases"]:
        mat = np.array(case["matrix"], dtype=np.float32)
        val, vec, rounds = sequential.max_eigen_value_and_vector(mat)
        assert rounds == case["rounds_main_py"], case["name"]
        if same_numpy:
            assert float(val) == case["eigen_val"] and [float(x) for x in vec] == case["eigen_vec"], case["name"]
        else:
            assert abs(float(val) - case["eigen_val"]) <= 1e-6 * case["eigen_val"]
            assert np.allclose(vec, case["eigen_vec"], rtol=1e-6, atol=0)


def test_sequential_model_against_the_unmodified_main_py_when_the_reference_tree_is_here():
    import importlib.util
    path = "/root/reference/main.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present (GPU box)")
    from oracle import sequential
    spec = importlib.util.spec_from_file_location("reference_main", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for mat in (oracle.hilbert(256), (oracle.uniform(100, 9) + np.float32(0.1)).astype(np.float32)):
        a = ref.max_eigen_value_and_vector(mat)
        b = sequential.max_eigen_value_and_vector(mat)
        assert a[2] == b[2] and a[0] == b[0] and np.array_equal(a[1], b[1])
