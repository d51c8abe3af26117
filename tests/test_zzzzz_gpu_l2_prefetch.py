"""L2 prefetch across the round barrier (st_options.kernel 21 / 22 / 23) and of the next unit during the pass
(24 / 25 / 26, static scheduling), DESIGN.md section 8.

Kernels 21-23 = resident-e configuration 13 plus one bulk L2 prefetch per warp (cp.async.bulk.prefetch.L2, SASS
UBLKPF.L2) of the unit the warp takes second in the next round, issued right before the round barrier.  A hint
cannot change a bit; this file holds it to that.

STATUS: written after round 1's GPU budget was spent; bit-identical on the emulated library (where the hint only
checks its address range).  First run on hardware at the end of round 1: the file sorts after every other GPU test
so that nothing it might do can disturb them.
"""
import numpy as np
import pytest

import oracle
from eigen_value_b200 import STOP_RELATIVE

pytestmark = pytest.mark.gpu


def _oracle(mat, **kw):
    return oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, **kw)


def _assert_same_bits(got, want, what):
    val, vec, it = got
    w_val, w_vec, _, w_it = want
    assert it == w_it, (what, it, w_it)
    assert np.float32(val).view(np.uint32) == np.float32(w_val).view(np.uint32), what
    assert np.array_equal(vec.view(np.uint32), w_vec.view(np.uint32)), what


def _matrix(dim):
    return (oracle.uniform(dim, 1000 + dim) + np.float32(0.25)).astype(np.float32)


@pytest.mark.parametrize("dim", [2052, 4100, 8200])
def test_l2_prefetch_variants_are_hints_only(solver, dim):
    """Kernels 21-23 = resident-e configuration 13 plus a bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the unit
    every warp takes second in the next round, issued before the round barrier.  A hint cannot change a bit:
    alternating / forward / static sweeps, rows of one and of two work units (8200: the second unit is 8 columns)."""
    mat = _matrix(dim)
    cap = 4
    want = _oracle(mat, max_itr=cap)
    d = solver.upload(mat)
    for kid in (21, 22, 23, 24, 25, 26):
        # st_options.sweep: bit 0 alternate the direction, bit 1 force static, bit 2 force dynamic scheduling (the
        # default is dynamic from N = 8192 up); kernels 24-26 keep two grabs in flight under dynamic scheduling
        for sweep in (1, 0, 3, 2, 5, 4):
            info, vec = solver.solve_device(d, dim, kernel=kid, max_iter=cap, sweep=sweep)
            assert info.kernel_id == kid
            _assert_same_bits((info.eigen_val, vec, info.iter_count), want, f"kernel {kid} sweep {sweep}")
    with pytest.raises(Exception):
        solver.solve_device(d, dim, kernel=22, stop=STOP_RELATIVE)          # tuning variants: absolute stop only


def test_l2_prefetch_variants_full_size_headline(solver):
    # Hilbert 8192 (README.md:76: 17 rounds): same bits as the default configuration
    dim = 8192
    d = solver.hilbert(dim)
    base, base_vec = solver.solve_device(d, dim)
    for kid, sweep in ((21, 1), (22, 1), (23, 1), (24, 3), (25, 3), (26, 3)):
        info, vec = solver.solve_device(d, dim, kernel=kid, sweep=sweep)
        assert info.kernel_id == kid and info.iter_count == base.iter_count == 17
        assert info.eigen_val == base.eigen_val and np.array_equal(vec, base_vec)
