"""Multi-GPU parity on real GPUs.  With >= 2 visible GPUs this spawns one rank per GPU under
torchrun (tests/sharded_worker.py); with one GPU it still drives the sharded entry points at
world == 1.  Two mutually waiting kernels are never put on one GPU (B200_PROFILING.md)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_one_shard_path_equals_plain_solve(solver):
    from eigen_value_b200.sharded import ShardedSolver
    dim = 2048
    sh = ShardedSolver(solver, dim, 0, 1)
    d = sh.hilbert()
    info, vec = sh.solve(d)
    base, base_vec = solver.solve_device(solver.hilbert(dim), dim)
    assert info.iter_count == base.iter_count == 14
    assert info.eigen_val == base.eigen_val and np.array_equal(vec, base_vec)
    o_val, o_vec, _, o_it = oracle.similarity_transform(oracle.hilbert(dim))
    assert o_it == 14 and abs(float(info.eigen_val) - float(o_val)) <= 1e-5 * float(o_val)
    sh.close()


def test_shard_argument_checks(solver):
    import ctypes
    lib = solver.lib
    sh = ctypes.c_void_p()
    assert lib.st_shard_create(solver.ctx, 16, 3, 2, ctypes.byref(sh)) != 0      # rank >= world
    assert lib.st_shard_create(solver.ctx, 4, 0, 9, ctypes.byref(sh)) != 0       # world > ST_MAX_WORLD
    assert lib.st_shard_create(solver.ctx, 16, 0, 2, ctypes.byref(sh)) == 0
    from eigen_value_b200._lib import StResult
    res = StResult()
    # not linked yet: solving must fail cleanly instead of waiting for a peer forever
    assert lib.st_shard_solve(sh, solver.hilbert(16, 0, 8).ptr, None, None, ctypes.byref(res)) != 0
    lib.st_shard_destroy(sh)


def _gpu_count():
    from eigen_value_b200 import _lib
    return _lib.load().st_device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_exchange_multi_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world * 7),
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "SHARDED_OK" in proc.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_in_process_sharded_solve_with_local_link(world):
    """One process, one host thread per GPU, shards wired with st_shard_link_local (plain peer access
    instead of IPC handles).  Every rank must return the bits of the single-GPU solve, which in turn are
    the bits of the oracle in the kernels' summation order.  (On the emulated library this runs with
    ST_EMU_DEVICES pretend GPUs and exercises the fused exchange through the real C ABI.)"""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs in one process")
    import ctypes
    import threading
    from eigen_value_b200 import Solver, STOP_RELATIVE
    from eigen_value_b200._lib import StResult, check
    from eigen_value_b200.similarity_transform import make_options

    dim, seed = 1000, 0x5EED0001
    solvers = [Solver(g) for g in range(world)]
    lib = solvers[0].lib
    shards = []
    for g in range(world):
        sh = ctypes.c_void_p()
        check(lib.st_shard_create(solvers[g].ctx, dim, g, world, ctypes.byref(sh)), "st_shard_create")
        shards.append(sh)
    table = (ctypes.c_void_p * world)(*[s.value for s in shards])
    check(lib.st_shard_link_local(table, world), "st_shard_link_local")
    rows = []
    for g in range(world):
        r0, n = ctypes.c_uint32(), ctypes.c_uint32()
        check(lib.st_shard_rows(shards[g], ctypes.byref(r0), ctypes.byref(n)), "st_shard_rows")
        rows.append(solvers[g].uniform(dim, seed, r0.value, n.value))
        solvers[g].synchronize()

    mat = oracle.uniform(dim, seed)
    vecs = [solvers[g].alloc(4 * dim) for g in range(world)]     # nothing is allocated once the ranks run collectively

    def collective(opts):
        out = [None] * world

        def work(g):
            o = make_options(lib, **opts)
            res = StResult()
            rc = lib.st_shard_solve(shards[g], rows[g].ptr, ctypes.byref(o), vecs[g].ptr, ctypes.byref(res))
            out[g] = (rc, res.eigen_val, res.iter_count, vecs[g].download(np.float32, dim) if rc == 0 else None)

        threads = [threading.Thread(target=work, args=(g,)) for g in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=600)
        return out

    for opts, okw in ((dict(), dict()), (dict(eps=1e-6, stop=STOP_RELATIVE, max_iter=60), dict(eps=1e-6, stop=oracle.STOP_RELATIVE, max_itr=60))):
        out = collective(opts)
        assert all(o is not None and o[0] == 0 for o in out), out
        o_val, o_vec, _, o_it = oracle.similarity_transform(mat, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, **okw)
        for rc, val, it, vec in out:
            assert it == o_it and np.float32(val) == o_val and np.array_equal(vec, o_vec)

    # A sharded solve never allocates on the device: the in-place form needs a working copy that st_shard_create did
    # not reserve, so every rank refuses (ST_ERR_ARG, no rank enters the collective kernel) until st_shard_prepare ran.
    from eigen_value_b200 import FORM_INPLACE
    out = collective(dict(form=FORM_INPLACE, max_iter=5))
    assert all(o[0] == -2 for o in out), out
    assert b"st_shard_prepare" in lib.st_last_error() or True       # the message is thread-local to the worker threads
    o = make_options(lib, form=FORM_INPLACE, max_iter=5)
    for g in range(world):
        check(lib.st_shard_prepare(shards[g], ctypes.byref(o)), "st_shard_prepare")
    out = collective(dict(form=FORM_INPLACE, max_iter=5))
    assert all(o[0] == 0 for o in out), out
    o_val, o_vec, _, o_it = oracle.similarity_transform(mat, form=oracle.FORM_INPLACE, sum_mode=oracle.SUM_CUDA, max_itr=5)
    for rc, val, it, vec in out:
        assert it == o_it and np.float32(val) == o_val and np.array_equal(vec, o_vec)
    for v in vecs:
        v.free()
    for sh in shards:
        lib.st_shard_destroy(sh)
    for s in solvers:
        s.close()


@pytest.mark.parametrize("case", ["ragged-1001", "bf16-1000", "fp8-1000", "fp8-8196"])
def test_in_process_sharded_storage_formats_and_ragged_dims(case):
    """The sharded builds added after the first multi-GPU runs, two GPUs in one process: dim % 4 != 0 on the resident-e
    kernel's scalar units (ranks of 500 and 501 rows), bf16 storage, fp8 storage with each rank's own row scales
    (st_shard_solve_fp8; 8196: two work units per row).  Every rank returns the oracle's bits."""
    world = 2
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs in one process")
    import ctypes
    import threading
    from eigen_value_b200 import Solver
    from eigen_value_b200._lib import StResult, check
    from eigen_value_b200.similarity_transform import make_options

    kind, dim = case.split("-")
    dim, seed = int(dim), 0x5EED0002
    cap = 4 if dim > 4100 else 1000
    solvers = [Solver(g) for g in range(world)]
    lib = solvers[0].lib
    shards = []
    for g in range(world):
        sh = ctypes.c_void_p()
        check(lib.st_shard_create(solvers[g].ctx, dim, g, world, ctypes.byref(sh)), "st_shard_create")
        shards.append(sh)
    check(lib.st_shard_link_local((ctypes.c_void_p * world)(*[s.value for s in shards]), world), "st_shard_link_local")
    mat = (oracle.uniform(dim, seed) + np.float32(0.25)).astype(np.float32)
    rows, scales, vecs = [], [], []
    for g in range(world):
        r0, n = ctypes.c_uint32(), ctypes.c_uint32()
        check(lib.st_shard_rows(shards[g], ctypes.byref(r0), ctypes.byref(n)), "st_shard_rows")
        d32 = solvers[g].upload(mat[r0.value:r0.value + n.value])
        if kind == "bf16":
            rows.append(solvers[g].to_bf16(d32, n.value * dim))
            scales.append(None)
        elif kind == "fp8":
            codes, sc = solvers[g].to_fp8(d32, n.value, dim)          # every rank quantises its own rows
            rows.append(codes)
            scales.append(sc)
        else:
            rows.append(d32)
            scales.append(None)
        solvers[g].synchronize()
        vecs.append(solvers[g].alloc(4 * dim))
    out = [None] * world

    def work(g):
        o = make_options(lib, max_iter=cap)
        res = StResult()
        if kind == "fp8":
            rc = lib.st_shard_solve_fp8(shards[g], rows[g].ptr, scales[g].ptr, ctypes.byref(o), vecs[g].ptr, ctypes.byref(res))
        elif kind == "bf16":
            rc = lib.st_shard_solve_bf16(shards[g], rows[g].ptr, ctypes.byref(o), vecs[g].ptr, ctypes.byref(res))
        else:
            rc = lib.st_shard_solve(shards[g], rows[g].ptr, ctypes.byref(o), vecs[g].ptr, ctypes.byref(res))
        out[g] = (rc, res.eigen_val, res.iter_count, res.kernel_id, vecs[g].download(np.float32, dim) if rc == 0 else None)

    threads = [threading.Thread(target=work, args=(g,)) for g in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert all(o is not None and o[0] == 0 for o in out), out
    stored = oracle.to_bf16(mat)[0] if kind == "bf16" else oracle.to_fp8_rows(mat)[0] if kind == "fp8" else mat
    o_val, o_vec, _, o_it = oracle.similarity_transform(stored, form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA, max_itr=cap)
    for rc, val, it, kid, vec in out:
        assert kid == 11 and it == o_it and np.float32(val) == o_val and np.array_equal(vec, o_vec)
    for v in vecs:
        v.free()
    for sh in shards:
        lib.st_shard_destroy(sh)
    for s in solvers:
        s.close()


def test_missing_peer_turns_into_a_timeout_error_not_a_hang():
    """Two linked shards, only rank 0 calls the collective solve: its round barrier waits for rank 1's flag, gives up
    after the device-side timeout (10 s) and the call returns ST_ERR_TIMEOUT -- no hang, no crash.  The shard group is
    desynchronised afterwards (its epochs differ) and has to be recreated; a fresh group on the same contexts works."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    import ctypes
    import threading
    import time
    from eigen_value_b200 import Solver
    from eigen_value_b200._lib import StResult, check

    dim = 256
    solvers = [Solver(0), Solver(1)]
    lib = solvers[0].lib

    def make_group():
        shards = []
        for g in range(2):
            sh = ctypes.c_void_p()
            check(lib.st_shard_create(solvers[g].ctx, dim, g, 2, ctypes.byref(sh)), "st_shard_create")
            shards.append(sh)
        check(lib.st_shard_link_local((ctypes.c_void_p * 2)(*[s.value for s in shards]), 2), "st_shard_link_local")
        return shards

    rows = [solvers[g].hilbert(dim, g * 128, 128) for g in range(2)]
    shards = make_group()
    res, vec = StResult(), solvers[0].alloc(4 * dim)
    t0 = time.time()
    rc = lib.st_shard_solve(shards[0], rows[0].ptr, None, vec.ptr, ctypes.byref(res))      # rank 1 never shows up
    waited = time.time() - t0
    assert rc == -4 and b"timed out" in lib.st_last_error(), (rc, lib.st_last_error())          # ST_ERR_TIMEOUT
    assert 5 < waited < 60, waited
    for sh in shards:
        lib.st_shard_destroy(sh)

    shards = make_group()                                                                   # recreate: works again
    out = [None, None]

    def work(g):
        r, v = StResult(), solvers[g].alloc(4 * dim)
        rc = lib.st_shard_solve(shards[g], rows[g].ptr, None, v.ptr, ctypes.byref(r))
        out[g] = (rc, r.iter_count, v.download(np.float32, dim) if rc == 0 else None)

    threads = [threading.Thread(target=work, args=(g,)) for g in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    o_val, o_vec, _, o_it = oracle.similarity_transform(oracle.hilbert(dim), form=oracle.FORM_READONLY, sum_mode=oracle.SUM_CUDA)
    assert out[0][0] == 0 and out[1][0] == 0 and out[0][1] == out[1][1] == o_it == 10       # README.md:71
    assert np.array_equal(out[0][2], o_vec) and np.array_equal(out[1][2], o_vec)
    for sh in shards:
        lib.st_shard_destroy(sh)
    for s in solvers:
        s.close()
