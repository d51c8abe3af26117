"""Property test of the launch planner through the C ABI: random st_options (kernel id, CTA size, grid size, sweep /
scheduling bits, form, stop test, accumulator, storage) on random small matrices.  Every call must either be
refused cleanly (a documented unsupported combination: StError, handle still usable) or return the oracle's bits.
Runs on the B200 under `-m gpu` and, inside the CPU suite, on the emulated library with the B200's launch shapes."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from eigen_value_b200 import ACC_F64, FORM_INPLACE, FORM_READONLY, STOP_RELATIVE
from eigen_value_b200._lib import StError

import os

pytestmark = pytest.mark.gpu


@st.composite
def cases(draw):
    # mostly supported combinations (so that the planner and the kernels are what is exercised), some not
    dim = draw(st.one_of(st.integers(1, 150).map(lambda k: 8 * k), st.integers(1, 300).map(lambda k: 4 * k), st.integers(1, 1200)))
    kernel = draw(st.sampled_from([0, 0, 0, 1, 1, 2, 2, 10, 11, 12, 13, 13, 17, 20, 24]))   # 17, 24: ids that no longer exist -> refused
    wild = draw(st.integers(0, 9)) == 0
    bf16 = draw(st.booleans()) if (wild or (kernel in (0, 1, 11) and dim % 8 == 0)) else False
    acc64 = draw(st.booleans()) if (wild or (not bf16 and kernel in (0, 1, 10, 12, 13))) else False
    form = draw(st.sampled_from([FORM_READONLY, FORM_INPLACE])) if (wild or (kernel in (0, 1) and not bf16 and not acc64)) else FORM_READONLY
    stop = draw(st.integers(0, 1)) if (wild or kernel in (0, 1, 2, 10, 11, 12, 13, 20)) else 0
    return dict(dim=dim, kernel=kernel, bf16=bf16, acc64=acc64, form=form, stop=stop,
                threads=draw(st.sampled_from([0, 0, 64, 128, 250, 256, 512, 1024])), ctas=draw(st.sampled_from([0, 0, 1, 3, 17, 148, 400])),
                sweep=draw(st.sampled_from([0, 1, 1, 3, 5])), max_iter=draw(st.sampled_from([1, 3, 25])),
                seed=draw(st.integers(0, 99)), hilbert=draw(st.booleans()))


@settings(max_examples=200, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_every_option_combination_is_refused_cleanly_or_bit_exact(solver, c):
    dim = c["dim"]
    mat = oracle.hilbert(dim) if c["hilbert"] else (oracle.uniform(dim, c["seed"]) + np.float32(0.25)).astype(np.float32)
    sum_mode, data = oracle.SUM_CUDA, mat
    if c["bf16"]:
        mat, data = oracle.to_bf16(mat)
        sum_mode = oracle.SUM_CUDA_BF16
    elif c["acc64"]:
        sum_mode = oracle.SUM_CUDA_F64
    kw = dict(kernel=c["kernel"], threads=c["threads"], ctas=c["ctas"], sweep=c["sweep"], form=c["form"], max_iter=c["max_iter"],
              stop=STOP_RELATIVE if c["stop"] else 0, accumulate=ACC_F64 if c["acc64"] else 0, bf16=c["bf16"])
    d = solver.upload(data)
    try:
        info, vec = solver.solve_device(d, dim, **kw)
    except StError as exc:
        assert "failed with code -2" in str(exc), (c, str(exc))         # ST_ERR_ARG: an unsupported combination, said so
        probe, _ = solver.solve_device(solver.upload(np.array([[2.0, 1.0], [1.0, 3.0]], dtype=np.float32)), 2, max_iter=3)
        assert probe.iter_count == 3                                    # the handle is still usable
        return
    finally:
        d.free()
    want = oracle.similarity_transform(mat, form=oracle.FORM_INPLACE if c["form"] == FORM_INPLACE else oracle.FORM_READONLY,
                                       sum_mode=sum_mode, max_itr=c["max_iter"], stop=c["stop"])
    assert info.iter_count == want[3], (c, info.iter_count, want[3])
    assert np.float32(info.eigen_val).view(np.uint32) == np.float32(want[0]).view(np.uint32), (c, float(info.eigen_val), float(want[0]))
    assert np.array_equal(vec.view(np.uint32), want[1].view(np.uint32)), c
