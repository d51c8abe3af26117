"""Fused peer-store exchange vs the collective baseline, same problem, same GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29555 tools/compare_exchange.py [N]

fused      : ShardedSolver.solve -- one persistent kernel per GPU, row sums stored into the peers'
             buffers from inside the kernel, one flag exchange per round (the product path)
collective : collective_round_loop -- host-driven rounds: row-pass kernel, torch.distributed
             all_gather_into_tensor (NCCL), max / stop / update kernels, one host read per round
             (what "call NCCL between kernels" costs; also how the reference drives its loop,
             similarity_transform.cpp:39-53)
Both must return the same eigenpair and round count.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from eigen_value_b200 import Solver  # noqa: E402
from eigen_value_b200.sharded import CudaRoundBackend, ShardedSolver, collective_round_loop  # noqa: E402


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = Solver(local)
    sh = ShardedSolver(solver, dim, rank, world)
    d_rows = sh.hilbert()
    d_vec = solver.alloc(4 * dim)

    def sync():
        dist.barrier()
        torch.cuda.synchronize()
        solver.synchronize()

    # fused
    for _ in range(3):
        sync()
        info, _ = sh.solve(d_rows, d_eigen_vec=d_vec)
    fused_ms = []
    for _ in range(5):
        sync()
        t0 = time.perf_counter()
        info, _ = sh.solve(d_rows, d_eigen_vec=d_vec)
        sync()
        fused_ms.append((time.perf_counter() - t0) * 1e3)
    fused_vec = d_vec.download(np.float32, dim)

    # collective
    backend = CudaRoundBackend(solver, d_rows, dim, sh.row0, sh.rows)
    dev = torch.device("cuda", local)
    coll_ms = []
    for i in range(3):
        sync()
        t0 = time.perf_counter()
        lam, e, it = collective_round_loop(backend, dim, rank, world, device=dev)
        sync()
        if i:
            coll_ms.append((time.perf_counter() - t0) * 1e3)
    assert it == info.iter_count, (it, info.iter_count)
    assert abs(lam - float(info.eigen_val)) <= 1e-6 * abs(lam)
    assert np.max(np.abs(e.cpu().numpy() - fused_vec)) <= 1e-6

    t = torch.tensor([min(fused_ms), min(coll_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        passes = info.passes
        print(json.dumps({
            "workload": f"hilbert-{dim}", "n_gpus": world, "rounds": info.iter_count,
            "fused_ms_wall": round(float(t[0]), 3), "fused_ms_device": round(info.loop_ms, 3),
            "fused_us_per_round": round(float(t[0]) * 1e3 / passes, 1),
            "collective_ms_wall": round(float(t[1]), 3),
            "collective_us_per_round": round(float(t[1]) * 1e3 / passes, 1),
            "speedup_fused_over_collective": round(float(t[1]) / float(t[0]), 2),
            "same_eigenpair_and_rounds": True}), flush=True)
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
