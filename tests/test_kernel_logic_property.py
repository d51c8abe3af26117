"""Property test of the round kernels on the CPU emulation harness: random launch shapes, sizes, GPU counts and
option combinations; the result must always be the oracle's, bit for bit (tests/cuda_emu, oracle.SUM_CUDA*).
Looks for corner cases of the index arithmetic: fewer rows than warps, row counts that do not divide, ragged
dimensions, more CTAs than rows, GPU counts that do not divide the matrix, every option at once."""
import os
import sys

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import emu  # noqa: E402


@st.composite
def cases(draw):
    kernel = draw(st.sampled_from([1, 1, 13, 10, 12, 11, 11, 20, 2, 2]))
    dim = draw(st.integers(1, 1100))
    if kernel not in (1, 11):
        dim = max(4, dim - dim % 4)                       # vector kernels: dim % 4 == 0 (11 also runs on scalar units)
    if kernel == 20:
        dim = min(dim, 512)
    world = 1 if kernel == 20 else draw(st.integers(1, min(8, dim)))
    storage = draw(st.sampled_from(["f32", "f32", "bf16", "fp8"])) if (kernel in (1, 11) and dim % 4 == 0) else "f32"
    bf16, fp8 = storage == "bf16", storage == "fp8"
    acc64 = storage == "f32" and kernel in (1, 13, 10, 12) and draw(st.booleans())
    form = draw(st.integers(0, 1)) if (kernel == 1 and storage == "f32" and not acc64) else 0
    stop = draw(st.integers(0, 1))
    return dict(kernel=kernel, dim=dim, world=world, bf16=bf16, fp8=fp8, acc64=acc64, form=form, stop=stop,
                threads=draw(st.sampled_from([32, 64, 96, 128, 256, 512])), ctas=draw(st.integers(1, 9)),
                dynamic=draw(st.integers(0, 1)), sweep=draw(st.integers(0, 1)),
                eps=draw(st.sampled_from([1e-3, 1e-5])), max_iter=draw(st.sampled_from([1, 2, 7, 40])),
                seed=draw(st.integers(0, 1000)), hilbert=draw(st.booleans()))


@settings(max_examples=600, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(cases())
def test_any_shape_any_option_matches_the_oracle(c):
    dim = c["dim"]
    mat = oracle.hilbert(dim) if c["hilbert"] else (oracle.uniform(dim, c["seed"]) + np.float32(0.25)).astype(np.float32)
    sum_mode, data, scale = oracle.SUM_CUDA, mat, None
    if c["bf16"]:
        mat, data = oracle.to_bf16(mat)                   # 4-element words in the fp32 kernels' order: SUM_CUDA
    elif c["fp8"]:
        mat, data, scale = oracle.to_fp8_rows(mat)        # the same, on the dequantised matrix
    elif c["acc64"]:
        sum_mode = oracle.SUM_CUDA_F64
    got = emu.solve(data, dim, kernel=c["kernel"], threads=c["threads"], ctas=c["ctas"], world=c["world"], form=c["form"],
                    stop=c["stop"], dynamic=c["dynamic"], sweep=c["sweep"], eps=c["eps"], max_iter=c["max_iter"],
                    bf16=c["bf16"], acc64=c["acc64"], fp8_scale=scale)
    want = oracle.similarity_transform(mat, form=oracle.FORM_INPLACE if c["form"] else oracle.FORM_READONLY, sum_mode=sum_mode,
                                       eps=c["eps"], max_itr=c["max_iter"], stop=c["stop"], ranks=c["world"])
    val, vec, it, passes, agree = got
    assert agree, c
    assert it == want[3], (c, it, want[3])
    assert np.float32(val).view(np.uint32) == np.float32(want[0]).view(np.uint32), (c, float(val), float(want[0]))
    assert np.array_equal(vec.view(np.uint32), want[1].view(np.uint32)), c
