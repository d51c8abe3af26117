"""Device group: several GPUs behind ONE handle (st_group_attach / ST_DEVICES), SURVEY 8(b) + 8(e).

The reference's wrapper knows one queue.  With a group attached, the same two symbols -- make_queue,
max_eigen_value -- shard large matrices row-block-wise over every GPU of the box: one host thread per GPU
uploads its row block and enters the collective round kernel.  Contract: the bits of the one-GPU solve.
Needs >= 2 GPUs in one process (skipped on a single-GPU box; runs on 4 pretend GPUs on the emulated library).

STATUS: written after round 1's GPU budget was spent; not yet run on multi-GPU hardware.
"""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import oracle
from eigen_value_b200 import EigenValue, Solver, STOP_RELATIVE, _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    return int(_lib.load().st_device_count())


def _want(mat, **kw):
    return oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, **kw)


def _same(got, want):
    val, vec, it = got
    w_val, w_vec, _, w_it = want
    assert it == w_it and np.float32(val) == np.float32(w_val) and np.array_equal(vec, w_vec)


@pytest.mark.parametrize("helpers", [1, 3, 7])
def test_drop_in_call_on_a_group_has_the_bits_of_one_gpu(helpers):
    if _gpus() < helpers + 1:
        pytest.skip(f"needs {helpers + 1} GPUs in one process")
    ev = EigenValue(devices=list(range(1, helpers + 1)), min_dim=256)
    assert ev.device_count == helpers + 1
    for kind, dim in (("hilbert", 1000), ("uniform", 1024), ("hilbert", 257)):
        mat = oracle.hilbert(dim) if kind == "hilbert" else oracle.uniform(dim, 0x5EED0001)
        keep = mat.copy()
        val, vec, ms, it = ev.similarity_transform(mat)
        assert np.array_equal(mat, keep) and ms >= 0
        _same((val, vec, it), _want(mat))
    # below min_dim the handle's own GPU does the work, same answer as ever
    val, vec, ms, it = ev.similarity_transform(oracle.hilbert(128))
    assert it == 9                                                            # reference README.md:70
    # two dimensions back to back, and the same one again: the exchange blocks are rebuilt per dimension
    for dim in (512, 300, 512):
        mat = oracle.hilbert(dim)
        val, vec, ms, it = ev.similarity_transform(mat)
        _same((val, vec, it), _want(mat))
    ev.so_lib.st_destroy(ev.sycl_q)


def test_solver_group_solve_host_with_options_and_detach():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs in one process")
    s = Solver(0)
    n = s.attach_group("all", min_dim=64)
    assert n == _gpus()
    mat = oracle.uniform(1000, 0x5EED0001)
    info, vec = s.solve_host(mat, eps=1e-6, stop=STOP_RELATIVE, max_iter=60)
    _same((info.eigen_val, vec, info.iter_count), _want(mat, eps=1e-6, stop=oracle.STOP_RELATIVE, max_itr=60))
    assert info.launches == min(n, 1000) and info.bytes_per_round == 4 * 1000 * 1000      # one launch per GPU, whole matrix
    # more devices than rows: the world shrinks to dim
    tiny = oracle.hilbert(64)[:3, :3].copy() + np.float32(1)
    lib = s.lib
    assert lib.st_group_detach(s.ctx) == 0 and lib.st_group_size(s.ctx) == 1
    s.attach_group("all", min_dim=1)
    info, vec = s.solve_host(tiny)
    _same((info.eigen_val, vec, info.iter_count), _want(tiny))
    # attaching twice, or naming the context's own device, is refused
    assert lib.st_group_attach(s.ctx, None, 0, 0) == -2
    s.detach_group()
    own = (ctypes.c_int * 1)(0)
    assert lib.st_group_attach(s.ctx, own, 1, 0) == -2
    info, vec = s.solve_host(mat, max_iter=5)                                  # no group: still works
    assert info.launches == 1
    s.close()


def test_concurrent_callers_on_a_group_handle():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs in one process")
    ev = EigenValue(devices="all", min_dim=128)
    mats = [oracle.hilbert(n) for n in (256, 384, 100, 512)]
    want = [_want(m) for m in mats]
    out = [None] * len(mats)

    def work(i):
        for _ in range(2):
            out[i] = ev.similarity_transform(mats[i])

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(mats))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for got, w in zip(out, want):
        assert got is not None
        _same((got[0], got[1], got[3]), w)
    ev.so_lib.st_destroy(ev.sycl_q)


def test_the_environment_variable_reaches_the_unmodified_call_sequence():
    """ST_DEVICES=all: make_queue + max_eigen_value exactly as the reference wrapper issues them
    (similarity_transform.py:35-37,66-76), in a fresh process.  On hardware the default threshold (8192) is
    exercised; the emulated device gets a smaller one through ST_GROUP_MIN_DIM."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs in one process")
    emulated = os.environ.get("ST_EMULATED_LIB") == "1"
    preamble = ("import sys; sys.path.insert(0, 'tests/cuda_emu'); import build as b\n"
                "_lib._build.SO_PATH = b.build_library(); _lib._build.stale = lambda: False\n") if emulated else ""
    code = ("import ctypes, os, numpy as np, oracle\n"
            "from eigen_value_b200 import _lib\n" + preamble +
            "lib = _lib.load()\n"
            "q = ctypes.c_void_p(); lib.make_queue(ctypes.byref(q)); assert q.value\n"
            "assert lib.st_group_size(q) == lib.st_device_count() >= 2\n"
            "dim = int(os.environ['GROUP_TEST_DIM']); mat = oracle.hilbert(dim)\n"
            "val = np.empty(1, np.float32); vec = np.empty(dim, np.float32); it = np.zeros(1, np.uint)\n"
            "ms = lib.max_eigen_value(q, mat.ctypes.data, val.ctypes.data, vec.ctypes.data, dim, it.ctypes.data)\n"
            "w = oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA)\n"
            "assert ms >= 0 and int(it[0]) == w[3] and val[0] == w[0] and np.array_equal(vec, w[1])\n"
            "print('GROUP_ENV_OK', lib.st_group_size(q))\n")
    env = dict(os.environ, ST_DEVICES="all", GROUP_TEST_DIM="1024" if emulated else "8192", PYTHONPATH=ROOT)
    if emulated:
        env["ST_GROUP_MIN_DIM"] = "256"
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=600)
    assert proc.returncode == 0 and "GROUP_ENV_OK" in proc.stdout, proc.stdout[-3000:]
    # a malformed list leaves the handle NULL (the reference wrapper's failure convention) with a message
    bad = ("import ctypes\nfrom eigen_value_b200 import _lib\n" + preamble +
           "lib = _lib.load(); q = ctypes.c_void_p(); lib.make_queue(ctypes.byref(q))\n"
           "assert not q.value and b'ST_DEVICES' in lib.st_last_error(); print('BAD_LIST_OK')\n")
    proc = subprocess.run([sys.executable, "-c", bad], cwd=ROOT, env=dict(env, ST_DEVICES="0,zero"), stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0 and "BAD_LIST_OK" in proc.stdout, proc.stdout[-3000:]


def test_multi_threaded_upload_of_a_pageable_matrix_changes_nothing_but_the_transfer():
    """ST_UPLOAD_THREADS (opt-in): a pageable host matrix is staged by several host threads through pinned double
    buffers instead of by the driver.  Same bits; sizes that are not a multiple of the 4 MiB chunk; pinned sources
    keep the direct path.  Fresh process: the variable is read when the context is created."""
    emulated = os.environ.get("ST_EMULATED_LIB") == "1"
    preamble = ("import sys; sys.path.insert(0, 'tests/cuda_emu'); import build as b\n"
                "_lib._build.SO_PATH = b.build_library(); _lib._build.stale = lambda: False\n") if emulated else ""
    code = ("import numpy as np, oracle\n"
            "from eigen_value_b200 import _lib\n" + preamble +
            "from eigen_value_b200 import EigenValue, Solver\n"
            "ev = EigenValue(); staged = 0\n"
            "for dim in (3000, 3072, 2900):\n"           # 34.3 / 36 / 32.08 MiB: above the 32 MiB threshold
            "    mat = (oracle.uniform(dim, 77 + dim) + np.float32(0.5)).astype(np.float32)\n"
            "    keep = mat.copy()\n"
            "    val, vec, ms, it = ev.similarity_transform(mat)\n"
            "    w = oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA)\n"
            "    assert it == w[3] and val == w[0] and np.array_equal(vec, w[1]) and np.array_equal(mat, keep), dim\n"
            "    staged += mat.nbytes\n"
            "    assert ev.so_lib.st_staged_upload_bytes(ev.sycl_q) == staged, dim\n"
            "    with ev.pinned(mat):\n"
            "        got = ev.similarity_transform(mat)\n"
            "    assert got[0] == val and np.array_equal(got[1], vec)\n"
            "    assert ev.so_lib.st_staged_upload_bytes(ev.sycl_q) == staged        # pinned: the direct copy\n"
            "s = Solver(0)\n"
            "info, vec, plan = s.solve_streamed(oracle.hilbert(512), device_budget=4 * 64 * 512 * 4, block_rows=64)\n"
            "assert info.iter_count == 12 and s.lib.st_staged_upload_bytes(s.ctx) == 0      # blocks below 32 MiB: direct\n"
            "big = oracle.hilbert(6144)                                  # three blocks of 2048 rows = 48 MiB, two slots\n"
            "info, vec, plan = s.solve_streamed(big, device_budget=2 * 2048 * 6144 * 4 + 4096, block_rows=2048, max_iter=3)\n"
            "w = oracle.similarity_transform(big, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, max_itr=3)\n"
            "assert info.iter_count == w[3] and info.eigen_val == w[0] and np.array_equal(vec, w[1])\n"
            "assert plan['blocks'] == 3 and plan['slots'] == 2 and plan['h2d_bytes_total'] == 5 * 2048 * 6144 * 4\n"
            "assert s.lib.st_staged_upload_bytes(s.ctx) == plan['h2d_bytes_total'] > big.nbytes    # every block upload was staged\n"
            "print('UPLOAD_THREADS_OK')\n")
    env = dict(os.environ, ST_UPLOAD_THREADS="3", PYTHONPATH=ROOT)
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True, timeout=900)
    assert proc.returncode == 0 and "UPLOAD_THREADS_OK" in proc.stdout, proc.stdout[-3000:]
