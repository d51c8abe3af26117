"""torchrun worker for the multi-GPU parity check (one rank per GPU, NCCL for the handle
exchange only).  Launched by tests/test_gpu_sharded.py and usable by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/sharded_worker.py

Checks, for several sizes and both forms: every rank returns the same (lambda, e, rounds);
they equal the single-GPU solve of the same kernel BIT FOR BIT (each row is reduced by exactly
one warp in a fixed order, so the partition cannot change a row sum); and they agree with the
CPU oracle to BASELINE.json's tolerances.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from eigen_value_b200 import STOP_RELATIVE, Solver  # noqa: E402
from eigen_value_b200.sharded import CudaRoundBackend, ShardedSolver, collective_round_loop  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = Solver(local)
    report = []
    cases = [("hilbert", 1024, 0), ("hilbert", 1024, 1), ("uniform", 1000, 0), ("uniform", 2049, 0),
             ("hilbert", 8192, 0), ("hilbert", 16384, 0), ("uniform", 4096, 1)]
    for kind, dim, form in cases:
        sh = ShardedSolver(solver, dim, rank, world)
        d_rows = sh.hilbert() if kind == "hilbert" else sh.uniform(0x5EED0001)
        dist.barrier()
        info, vec = sh.solve(d_rows, form=form, max_iter=40)
        # single-GPU run of the same problem on this rank's GPU
        d_full = solver.hilbert(dim) if kind == "hilbert" else solver.uniform(dim, 0x5EED0001)
        one, one_vec = solver.solve_device(d_full, dim, form=form, max_iter=40)
        assert info.iter_count == one.iter_count, (kind, dim, form, info.iter_count, one.iter_count)
        assert info.eigen_val == one.eigen_val, (kind, dim, form, info.eigen_val, one.eigen_val)
        assert np.array_equal(vec, one_vec), (kind, dim, form)
        # all ranks identical
        t = torch.from_numpy(vec).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref)
        # oracle tolerance (skip the big ones on all but rank 0 to save host time)
        if dim <= 4096 or rank == 0:
            mat = oracle.hilbert(dim) if kind == "hilbert" else oracle.uniform(dim, 0x5EED0001)
            o_val, o_vec, _, o_it = oracle.similarity_transform(mat, max_itr=40, form=oracle.FORM_READONLY)
            assert abs(info.iter_count - o_it) <= (0 if kind == "hilbert" else 1)
            if info.iter_count == o_it:
                assert abs(float(info.eigen_val) - float(o_val)) <= 1e-5 * abs(float(o_val))
                assert np.max(np.abs(vec / vec.max() - o_vec / o_vec.max())) <= 1e-4
        report.append({"case": f"{kind}-{dim}-form{form}", "rounds": info.iter_count,
                       "lambda": float(info.eigen_val), "us_per_round": info.round_us_median,
                       "us_per_round_1gpu": one.round_us_median})
        d_rows.free()
        d_full.free()
        sh.close()

    # BASELINE configs 3 to 5 at full size, sharded: the wide kernel (N > 32768) and the resident-e kernel must return the
    # bits the CPU oracle computed matrix-free (tests/golden/generated_expected.json: lambda, rounds, eigenvector digest)
    import hashlib
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generated_expected.json")) as f:
        expected = json.load(f)["cases"]
    for name in ("hilbert-32768", "hilbert-65536", "uniform-65536", "hilbert-131072", "uniform-131072"):
        e = expected[name]
        dim = e["dim"]
        if 4 * dim * dim // world > 0.8 * solver.hbm_bytes:
            continue
        sh = ShardedSolver(solver, dim, rank, world)
        d_rows = sh.hilbert() if e["kind"] == "hilbert" else sh.uniform(e["seed"])
        dist.barrier()
        info, vec = sh.solve(d_rows, max_iter=e["max_iter"])
        assert info.iter_count == e["iter_count"], (name, info.iter_count)
        assert int(np.float32(info.eigen_val).view(np.uint32)) == e["eigen_val_bits"], (name, float(info.eigen_val))
        assert hashlib.sha256(np.ascontiguousarray(vec).tobytes()).hexdigest() == e["eigen_vec_sha256"], name
        report.append({"case": f"{name}-full-size-vs-cpu-oracle", "rounds": info.iter_count, "lambda": float(info.eigen_val),
                       "kernel": info.kernel_id, "us_per_round": info.round_us_median})
        d_rows.free()
        sh.close()

    # ST_STOP_RELATIVE (extension): the decision is taken redundantly on every rank from identical
    # inputs, so the sharded solve stops in the same round with the same bits as one GPU
    for kind, dim in (("uniform", 4096), ("hilbert", 16384)):
        sh = ShardedSolver(solver, dim, rank, world)
        d_rows = sh.hilbert() if kind == "hilbert" else sh.uniform(0x5EED0001)
        dist.barrier()
        info, vec = sh.solve(d_rows, eps=1e-6, stop=STOP_RELATIVE, max_iter=60)
        d_full = solver.hilbert(dim) if kind == "hilbert" else solver.uniform(dim, 0x5EED0001)
        one, one_vec = solver.solve_device(d_full, dim, eps=1e-6, stop=STOP_RELATIVE, max_iter=60)
        assert info.iter_count == one.iter_count and info.eigen_val == one.eigen_val, (kind, dim)
        assert np.array_equal(vec, one_vec), (kind, dim)
        report.append({"case": f"{kind}-{dim}-relative-stop", "rounds": info.iter_count, "lambda": float(info.eigen_val)})
        d_rows.free()
        d_full.free()
        sh.close()

    # the collective (all-gather) variant of the loop gives the same answer too
    dim = 1024
    sh = ShardedSolver(solver, dim, rank, world)
    d_rows = sh.hilbert()
    backend = CudaRoundBackend(solver, d_rows, dim, sh.row0, sh.rows)
    lam, e, it = collective_round_loop(backend, dim, rank, world, device=torch.device("cuda", local))
    info, vec = sh.solve(d_rows)
    assert it == info.iter_count == 13
    assert abs(lam - float(info.eigen_val)) <= 1e-6 * lam
    assert np.max(np.abs(e.cpu().numpy() - vec)) <= 1e-6
    report.append({"case": "collective-hilbert-1024", "rounds": it, "lambda": lam})
    sh.close()

    dist.barrier()
    if rank == 0:
        print("SHARDED_OK " + json.dumps(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
