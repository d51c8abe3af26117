"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same
inputs.  Tolerances are BASELINE.json's: eigenvalue 1e-5 relative, eigenvector 1e-4 max-abs
after normalisation, identical round count (the cases here sit away from the convergence
boundary; boundary cases are asserted as |delta| <= 1)."""
import numpy as np
import pytest

import oracle
from eigen_value_b200 import EigenValue, FORM_INPLACE, FORM_READONLY

pytestmark = pytest.mark.gpu

LAMBDA_RTOL = 1e-5
VEC_ATOL = 1e-4
HILBERT_ROUNDS = {128: 9, 256: 10, 512: 12, 1024: 13, 2048: 14, 4096: 15, 8192: 17}  # README.md:70-76
A3 = np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)


def normalised(v):
    return v / v.max()


def assert_parity(got, want, same_rounds=True):
    g_val, g_vec, g_it = got
    w_val, w_vec, w_it = want
    if same_rounds:
        assert g_it == w_it, (g_it, w_it)
    else:
        assert abs(g_it - w_it) <= 1, (g_it, w_it)
    if g_it == w_it:
        assert abs(float(g_val) - float(w_val)) <= LAMBDA_RTOL * abs(float(w_val)), (g_val, w_val)
        assert np.max(np.abs(normalised(g_vec) - normalised(w_vec))) <= VEC_ATOL


# ---- the drop-in boundary: make_queue + max_eigen_value ---------------------------------------
@pytest.fixture(scope="module")
def ev():
    return EigenValue()


def test_three_by_three_golden_through_the_boundary(ev):
    # reference tests/test.cpp:96-102
    val, vec, ms, it = ev.similarity_transform(A3)
    assert abs(val - 7.53114) < 1e-3
    assert abs(vec[0] - 0.394074) < 1e-3 and abs(vec[1] - 0.578844) < 1e-3 and abs(vec[2] - 0.997451) < 1e-3
    assert it == 4 and ms >= 0
    o_val, o_vec, _, o_it = oracle.similarity_transform(A3)
    assert_parity((val, vec, it), (o_val, o_vec, o_it))


def test_reference_acceptance_test(ev):
    # reference wrapper/python/test.py:8-18, verbatim criterion, 4 repeats on one handle
    mat = np.random.default_rng(1234).random((1 << 10, 1 << 10)).astype("f")
    keep = mat.copy()
    for _ in range(4):
        lam, v, ts, itr = ev.similarity_transform(mat)
        assert np.all(np.isclose(np.matmul(mat, v), lam * v, atol=1e-3)), "Av = lambda v assertion failed !"
        assert ts >= 0 and 1 <= itr < 10
    assert np.array_equal(mat, keep), "caller's matrix must not be modified (similarity_transform.cpp:14,19)"


def test_iter_count_slot_gets_exactly_four_bytes(ev):
    # the reference wrapper passes an 8-byte np.uint slot, C writes a 4-byte uint (similarity_transform.py:63-73)
    import ctypes
    lib = ev.so_lib
    slot = np.full(1, 0xAAAAAAAA_00000000, dtype=np.uint64)
    val = np.empty(1, np.float32)
    vec = np.empty(3, np.float32)
    rc = lib.max_eigen_value(ev.sycl_q, A3.ctypes.data, val.ctypes.data, vec.ctypes.data, 3, slot.ctypes.data)
    assert rc >= 0
    assert int(slot[0]) == 0xAAAAAAAA_00000004


@pytest.mark.parametrize("dim", sorted(HILBERT_ROUNDS))
def test_hilbert_sweep_matches_published_rounds_and_oracle(ev, dim):
    # BASELINE config 2 / reference main.cpp:23-35 + README.md:70-76
    H = oracle.hilbert(dim)
    val, vec, ms, it = ev.similarity_transform(H)
    assert it == HILBERT_ROUNDS[dim]
    o_val, o_vec, _, o_it = oracle.similarity_transform(H)
    assert_parity((val, vec, it), (o_val, o_vec, o_it))


@pytest.mark.parametrize("dim", [1, 2, 3, 5, 31, 33, 100, 257, 1000, 1023, 4100])
def test_ragged_and_tiny_dims(ev, dim):
    # the reference breaks on dim=1 (wg_size 0) and on dim % wg_size != 0
    # (wrapper/similarity_transform.cpp:33); the replacement accepts any dim >= 1
    mat = (oracle.uniform(dim, seed=dim) + np.float32(0.25)).astype(np.float32)
    val, vec, ms, it = ev.similarity_transform(mat)
    o_val, o_vec, _, o_it = oracle.similarity_transform(mat)
    assert_parity((val, vec, it), (o_val, o_vec, o_it), same_rounds=False)
    if it == o_it and dim > 1:
        assert np.all(np.isclose(mat @ vec, val * vec, atol=2e-3 * max(1.0, float(val))))


def test_bad_arguments_do_not_crash(ev):
    lib = ev.so_lib
    assert lib.max_eigen_value(ev.sycl_q, None, None, None, 4, None) < 0
    assert lib.max_eigen_value(None, A3.ctypes.data, A3.ctypes.data, A3.ctypes.data, 3, A3.ctypes.data) < 0
    assert lib.st_last_error()
    # empty input: dim == 0 is rejected, nothing is written
    val = np.full(1, -1.0, np.float32)
    slot = np.full(1, 77, np.uint32)
    assert lib.max_eigen_value(ev.sycl_q, A3.ctypes.data, val.ctypes.data, A3.ctypes.data, 0, slot.ctypes.data) < 0
    assert val[0] == -1.0 and slot[0] == 77


@pytest.mark.parametrize("dim", [3, 64, 1000])
def test_nan_input_never_converges_and_hits_the_cap(ev, dim):
    """A NaN anywhere makes every comparison of the stop test false, so the reference runs its
    full MAX_ITR = 1000 rounds (similarity_transform.cpp:39,54); so must we,
    and the handle has to stay usable afterwards."""
    mat = (oracle.uniform(dim, 5) + np.float32(0.5)).astype(np.float32)
    mat[dim // 2, dim // 3] = np.nan
    val, vec, ms, it = ev.similarity_transform(mat)
    assert it == 1000
    o_val, o_vec, _, o_it = oracle.similarity_transform(mat)
    assert o_it == 1000
    val, vec, ms, it = ev.similarity_transform(A3)          # still fine after the poisoned solve
    assert it == 4 and abs(val - 7.53114) < 1e-3


# ---- per-kernel entry points vs the reference's unit fixtures (tests/test.cpp:22-73) -----------
N = 1 << 10


def test_kernel_sum_across_rows_identity(solver):
    assert np.all(solver.sum_across_rows(np.eye(N, dtype=np.float32)) == 1.0)


def test_kernel_sum_across_rows_random_vs_oracle(solver):
    for dim in (7, 64, 1000, 2048):
        m = oracle.uniform(dim, seed=99 + dim)
        got, want = solver.sum_across_rows(m), oracle.sum_across_rows(m)
        assert np.max(np.abs(got - want) / want) < 2e-6


def test_kernel_find_max(solver):
    assert solver.find_max(np.arange(1, N + 1, dtype=np.float32)) == N
    assert solver.find_max(-np.ones(5, dtype=np.float32)) == 0.0   # zero-filled cell, :169


def test_kernel_compute_eigen_vector(solver):
    v = np.arange(1, N + 1, dtype=np.float32)
    e = solver.initialise_eigen_vector(N)
    assert np.all(e == 1.0)
    e = solver.compute_eigen_vector(v, float(N), e)
    want = np.ones(N, np.float32)
    oracle.compute_eigen_vector(v, float(N), want)
    assert np.array_equal(e, want)


def test_kernel_stop(solver):
    ok = np.full(N, np.float32(1.0) + np.float32(1e-4), dtype=np.float32)
    bad = (np.arange(1, N + 1, dtype=np.float32) * np.float32(1e-4)).astype(np.float32)
    assert solver.stop(ok) == 1 and oracle.stop(ok) == 1
    assert solver.stop(bad) == 0 and oracle.stop(bad) == 0          # fails through the wrap pair only
    edge = np.array([0.0, 1e-3, 0.0, 0.0], dtype=np.float32)
    assert solver.stop(edge) == oracle.stop(edge) == 0               # strict <


def test_kernel_compute_next_matrix_bit_exact(solver):
    for dim in (5, 64, 512):
        W = oracle.uniform(dim, seed=5 + dim) + np.float32(0.5)
        s = oracle.sum_across_rows(W)
        got = solver.compute_next_matrix(W, s)
        want = W.copy()
        oracle.compute_next_matrix(want, s)
        assert np.array_equal(got, want)


# ---- input generation on the device -------------------------------------------------------------
def test_device_hilbert_bit_exact(solver):
    for dim, row0, rows in ((64, 0, 64), (1000, 0, 1000), (1000, 333, 100), (4096, 4000, 96)):
        got = solver.hilbert(dim, row0, rows).download(np.float32, rows * dim).reshape(rows, dim)
        assert np.array_equal(got, oracle.hilbert(dim, row0, rows))


def test_device_uniform_bit_exact_and_shard_independent(solver):
    for dim, row0, rows in ((64, 0, 64), (37, 0, 37), (37, 5, 9), (1001, 17, 300), (2048, 1024, 512)):
        got = solver.uniform(dim, 0x5EED0001, row0, rows).download(np.float32, rows * dim).reshape(rows, dim)
        assert np.array_equal(got, oracle.uniform(dim, 0x5EED0001, row0, rows))


# ---- device-resident solves, both forms, options -------------------------------------------------
@pytest.mark.parametrize("form", [FORM_READONLY, FORM_INPLACE])
@pytest.mark.parametrize("dim", [3, 128, 1000, 2048])
def test_forms_agree_with_oracle(solver, form, dim):
    H = oracle.hilbert(dim) if dim > 3 else A3
    info, vec = solver.solve_device(solver.upload(H), dim, form=form)
    o_val, o_vec, _, o_it = oracle.similarity_transform(
        H, form=oracle.FORM_READONLY if form == FORM_READONLY else oracle.FORM_INPLACE)
    assert_parity((info.eigen_val, vec, info.iter_count), (o_val, o_vec, o_it))
    assert info.passes == info.iter_count + 1 and info.launches == 1


@pytest.mark.parametrize("threads", [256, 512, 1024])
@pytest.mark.parametrize("sweep", [0, 1])
def test_launch_shape_and_sweep_do_not_change_results(solver, threads, sweep):
    d = solver.hilbert(2048)
    base, base_vec = solver.solve_device(d, 2048)
    info, vec = solver.solve_device(d, 2048, threads=threads, sweep=sweep)
    assert info.iter_count == base.iter_count == 14
    assert info.eigen_val == base.eigen_val and np.array_equal(vec, base_vec)


@pytest.mark.parametrize("dim", [256, 4096, 8192, 16384])
def test_kernel_variants_are_bitwise_identical(solver, dim):
    """general chunked loop (1), resident-e kernel (13, 10, 12, 11), on-chip cluster kernel (20, N <= 512):
    one evaluation order, so every variant returns the same bits."""
    d = solver.uniform(dim, 0x5EED0000 + dim)
    base, base_vec = solver.solve_device(d, dim, kernel=1, max_iter=6)
    kids = (0, 13, 10, 11) + ((12,) if dim <= 8192 else ()) + ((20,) if dim <= 512 else ())   # 12: 3 x 4 KB slots per warp + e
    for kid in kids:
        info, vec = solver.solve_device(d, dim, kernel=kid, max_iter=6)
        assert info.iter_count == base.iter_count, kid
        assert info.eigen_val == base.eigen_val and np.array_equal(vec, base_vec), kid
        assert info.kernel_id == (kid if kid else (20 if dim <= 512 else 13))


@pytest.mark.parametrize("dim", [4, 8, 100, 128, 256, 384, 512])
def test_on_chip_cluster_kernel_matches_general_loop(solver, dim):
    """N <= 512: matrix resident in the shared memory of one cluster (1/2/4/8 CTAs), row sums
    exchanged through distributed shared memory.  Same bits as the general loop, same rounds as
    the oracle."""
    mat = oracle.hilbert(dim)
    d = solver.upload(mat)
    base, base_vec = solver.solve_device(d, dim, kernel=1)
    info, vec = solver.solve_device(d, dim, kernel=20)
    assert info.kernel_id == 20 and info.grid in (1, 2, 4, 8)
    assert info.iter_count == base.iter_count and info.eigen_val == base.eigen_val
    assert np.array_equal(vec, base_vec)
    o_val, o_vec, _, o_it = oracle.similarity_transform(mat)
    assert_parity((info.eigen_val, vec, info.iter_count), (o_val, o_vec, o_it))


def test_zero_copy_torch_tensor_solve(solver):
    # device-resident input without PCIe staging
    import torch
    H = torch.from_numpy(oracle.hilbert(1024)).cuda()
    info, vec = solver.solve_tensor(H)
    assert info.iter_count == 13 and vec.is_cuda and vec.shape == (1024,)
    o_val, o_vec, _, o_it = oracle.similarity_transform(oracle.hilbert(1024))
    assert_parity((info.eigen_val, vec.cpu().numpy(), info.iter_count), (o_val, o_vec, o_it))
    with pytest.raises(ValueError):
        solver.solve_tensor(H[:, :512])                 # not square
    with pytest.raises(ValueError):
        solver.solve_tensor(H.double())                 # not float32
    with pytest.raises(ValueError):
        solver.solve_tensor(H.t())                      # not C-contiguous


def test_many_rounds_beyond_the_stamp_buffer(solver):
    # max_iter far above the reference's 1000: rounds past the stamp buffer still run
    d = solver.upload(np.array([[2.0, 1.0], [1.0, 3.0]], dtype=np.float32))
    info, vec = solver.solve_device(d, 2, eps=0.0, max_iter=70000)    # eps 0 never converges
    assert info.iter_count == 70000 and info.passes == 70000
    lam = (5 + 5 ** 0.5) / 2
    assert abs(float(info.eigen_val) - lam) < 1e-5 * lam


def test_max_iter_cap_reports_cap(solver):
    # never-converging case: iter_count == max_iter (reference similarity_transform.cpp:39,54)
    d = solver.hilbert(512)
    info, vec = solver.solve_device(d, 512, max_iter=5)
    assert info.iter_count == 5 and info.passes == 5
    o_val, o_vec, _, o_it = oracle.similarity_transform(oracle.hilbert(512), max_itr=5)
    assert_parity((info.eigen_val, vec, info.iter_count), (o_val, o_vec, o_it))


def test_eps_option(solver):
    d = solver.hilbert(1024)
    info, vec = solver.solve_device(d, 1024, eps=1e-2)
    o_val, o_vec, _, o_it = oracle.similarity_transform(oracle.hilbert(1024), eps=1e-2)
    assert_parity((info.eigen_val, vec, info.iter_count), (o_val, o_vec, o_it))


def test_multi_chunk_columns_16384(solver):
    # N > 8192 columns exercises the chunked scale-vector staging; oracle takes ~2 s
    H = oracle.hilbert(16384)
    info, vec = solver.solve_device(solver.hilbert(16384), 16384)
    o_val, o_vec, _, o_it = oracle.similarity_transform(H, form=oracle.FORM_READONLY)
    assert info.iter_count == 18                                     # BASELINE.md section 5 prediction
    assert_parity((info.eigen_val, vec, info.iter_count), (o_val, o_vec, o_it))


def _row_residuals(kind, dim, seed, rows, lam, vec):
    """|(A v)[r] - lambda v[r]| / |lambda v[r]| for a few rows rebuilt on the host in fp64."""
    out = []
    v64 = vec.astype(np.float64)
    for r in rows:
        a = (oracle.hilbert(dim, r, 1) if kind == "hilbert" else oracle.uniform(dim, seed, r, 1))[0]
        lhs = float(a.astype(np.float64) @ v64)
        out.append(abs(lhs - float(lam) * float(vec[r])) / abs(float(lam) * float(vec[r])))
    return out


def test_full_size_hilbert_131072_on_one_gpu(solver):
    """BASELINE's largest matrix (64 GiB) on one B200, solved to convergence (~0.25 s).  The oracle
    cannot hold it, so the checks are size-independent: the round count and lambda predicted in
    BASELINE.md section 5 (23 rounds, 2.7381425), A.v ~= lambda.v on rows rebuilt on the host, and
    the stop criterion itself: the circular test is binding through the wrap pair, i.e.
    max(s) - min(s) < EPS, so every row's s = (A.v)[r]/v[r] must sit within EPS of lambda."""
    if solver.hbm_bytes < 80 * 2**30:
        pytest.skip("needs 64 GiB of device memory")
    dim = 131072
    d = solver.hilbert(dim)
    info, vec = solver.solve_device(d, dim)
    d.free()
    assert info.iter_count == 23 and info.passes == 24
    assert abs(float(info.eigen_val) - 2.7381425) <= 1e-5 * 2.7381425
    assert vec.min() > 0 and 0.99 < vec.max() <= 1.0
    res = _row_residuals("hilbert", dim, 0, [0, 1, 65535, 100000, dim - 1], info.eigen_val, vec)
    # the returned vector is e after this round's update, so the residual is the NEXT round's
    # s[r] against this round's s[0]: spread < EPS plus one round's drift of lambda (~3e-4)
    assert max(res) < 2e-3 / 2.7


def test_full_size_property_uniform_65536(solver):
    """BASELINE config 4's matrix (16 GiB, seed 0x5EED0001), capped at 10 rounds: after the
    first handful of rounds the iteration sits at its fp32 noise floor, so
    A.v ~= lambda.v must hold to ~1e-6 on rows rebuilt on the host, and lambda ~ N/2."""
    if solver.hbm_bytes < 40 * 2**30:
        pytest.skip("needs 16 GiB of device memory")
    dim = 65536
    d = solver.uniform(dim, 0x5EED0001)
    info, vec = solver.solve_device(d, dim, max_iter=10)
    d.free()
    assert info.iter_count == 10                                     # never converges in fp32
    assert abs(float(info.eigen_val) - dim / 2) < 0.01 * dim
    res = _row_residuals("uniform", dim, 0x5EED0001, [0, 7, 40000, dim - 1], info.eigen_val, vec)
    assert max(res) < 1e-5


def test_full_size_property_uniform_32768(solver):
    """BASELINE-size property check where the oracle is too slow: A.v ~= lambda.v on a random
    (0,1] 32768^2 matrix after a capped run, evaluated on the GPU-generated rows."""
    dim = 32768
    d = solver.uniform(dim, 0x5EED0001)
    info, vec = solver.solve_device(d, dim, max_iter=12)
    assert info.iter_count == 12                                     # never converges in fp32
    rows = [0, 1, 12345, dim - 1]
    for r in rows:
        a = oracle.uniform(dim, 0x5EED0001, row0=r, rows=1)[0].astype(np.float64)
        lhs = float(a @ vec.astype(np.float64))
        assert abs(lhs - float(info.eigen_val) * float(vec[r])) <= 1e-5 * abs(lhs)


# ---- the CUDA path against outputs of the reference itself (tests/golden/reference_sycl.json) ----
from golden_util import build_matrix as _golden_matrix, expected as _golden_expected, load_reference_cases  # noqa: E402

_GOLDEN_CASES = load_reference_cases()["cases"]


@pytest.mark.parametrize("case", _GOLDEN_CASES, ids=[c["name"] for c in _GOLDEN_CASES])
def test_cuda_matches_the_reference_own_outputs(ev, case):
    """Fixture = what the unmodified reference C++ returned on the CPU SYCL shim
    (tests/golden/make_reference_golden.py).  Through make_queue/max_eigen_value the CUDA path
    must agree within BASELINE.json's tolerances."""
    mat = _golden_matrix(case)
    r_val, r_vec, r_it = _golden_expected(case)
    val, vec, ms, it = ev.similarity_transform(mat)
    assert_parity((val, vec, it), (r_val, r_vec, r_it), same_rounds=case["kind"] != "uniform")


def test_pinning_the_callers_matrix_changes_nothing_but_the_transfer(ev):
    # st_pin_host / st_unpin_host (extension): same result, the matrix is left untouched, and usable again after unpinning
    mat = oracle.hilbert(1024)
    keep = mat.copy()
    base = ev.similarity_transform(mat)
    with ev.pinned(mat):
        for _ in range(2):
            got = ev.similarity_transform(mat)
            assert got[0] == base[0] and got[3] == base[3] == 13 and np.array_equal(got[1], base[1])
    assert np.array_equal(mat, keep)
    got = ev.similarity_transform(mat)
    assert got[0] == base[0]
    lib = ev.so_lib
    assert lib.st_pin_host(ev.sycl_q, None, 16) != 0 and lib.st_unpin_host(ev.sycl_q, None) != 0


def test_device_out_of_memory_is_reported_as_nomem_and_the_handle_survives(solver):
    # a matrix far beyond the device's memory (dim^2 * 4 B = 4 TiB): st_malloc must fail with ST_ERR_NOMEM, not crash,
    # and the context must stay usable (an allocation failure is not a sticky CUDA error)
    import ctypes
    lib = solver.lib
    p = ctypes.c_void_p()
    rc = lib.st_malloc(solver.ctx, 4 << 40, ctypes.byref(p))
    assert rc == -5 and b"memory" in lib.st_last_error().lower(), (rc, lib.st_last_error())      # ST_ERR_NOMEM
    info, _ = solver.solve_device(solver.hilbert(256), 256)
    assert info.iter_count == 10


def test_concurrent_callers_on_one_handle_are_serialised(ev):
    """ctypes releases the GIL, so Python threads can enter max_eigen_value on ONE handle at the same time (the
    reference is re-entrant by accident, SURVEY 8(b) "Threading").  The handle serialises them; every caller gets
    its own correct result."""
    import threading
    mats = [oracle.hilbert(n) for n in (128, 256, 384, 512, 640, 1000)]
    want = [oracle.similarity_transform(m, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA) for m in mats]
    out = [None] * len(mats)

    def work(i):
        for _ in range(3):
            out[i] = ev.similarity_transform(mats[i])

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(mats))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    for got, w in zip(out, want):
        assert got is not None and got[3] == w[3] and got[0] == w[0] and np.array_equal(got[1], w[1])


# ---- BASELINE configs 3 to 5 at FULL size on one GPU against CPU-computed expected bits ---------------------------------
@pytest.mark.parametrize("name", ["hilbert-16384", "hilbert-32768", "uniform-32768", "hilbert-65536", "uniform-65536",
                                  "hilbert-131072", "uniform-131072"])
def test_full_size_configs_return_the_bits_the_cpu_oracle_computed(solver, name):
    """tests/golden/generated_expected.json holds lambda bits, round count and the sha256 of the raw eigenvector that the
    oracle's matrix-free loop (ORACLE_SUM_CUDA order) computed on the CPU -- Hilbert to convergence, the uniform cases
    (seeds of BASELINE configs 4 and 5) capped at 50 rounds.  The matrix is generated on the device: 16 GiB at 65536,
    64 GiB at 131072 (the north-star size on ONE GPU, wide kernel)."""
    import hashlib
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "generated_expected.json")) as f:
        e = json.load(f)["cases"][name]
    dim = e["dim"]
    if 4 * dim * dim > 0.8 * solver.hbm_bytes:
        pytest.skip("matrix does not fit this device")
    d = solver.hilbert(dim) if e["kind"] == "hilbert" else solver.uniform(dim, e["seed"])
    try:
        info, vec = solver.solve_device(d, dim, max_iter=e["max_iter"])
    finally:
        d.free()
    assert info.kernel_id == (2 if dim > 32768 else 13)
    assert info.iter_count == e["iter_count"]
    assert int(np.float32(info.eigen_val).view(np.uint32)) == e["eigen_val_bits"], float(info.eigen_val)
    assert hashlib.sha256(np.ascontiguousarray(vec).tobytes()).hexdigest() == e["eigen_vec_sha256"]
