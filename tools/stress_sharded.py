#!/usr/bin/env python
"""Soak test of the fused cross-GPU exchange (flat round barrier whose counters and max slots are never reset, solves
starting on the buffer parity the previous one did not end on): thousands of sharded solves back to back on one shard group,
round caps chosen so that consecutive solves end on either parity, every result compared bit for bit with the first
solve of the same (size, cap) and across ranks.  One JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tools/stress_sharded.py [--seconds 40]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from eigen_value_b200 import Solver  # noqa: E402
from eigen_value_b200.sharded import ShardedSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=40.0)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = Solver(local)
    report = []
    for kind, dim in (("uniform", 256), ("hilbert", 1024), ("uniform", 4096), ("hilbert", 8192), ("hilbert", 40960)):
        sh = ShardedSolver(solver, dim, rank, world)
        d_rows = sh.hilbert() if kind == "hilbert" else sh.uniform(0x5EED0001)
        d_vec = solver.alloc(4 * dim)
        caps = (1, 2, 3, 4, 7, 1000)
        first = {}
        solves = 0
        budget = args.seconds / 5.0
        # every rank must run the same number of solves: rank 0's clock decides, one flag per lap
        stop = torch.zeros(1, dtype=torch.int32, device="cuda")
        t0 = time.time()
        while True:
            for cap in caps:
                info, _ = sh.solve(d_rows, d_eigen_vec=d_vec, max_iter=cap)
                vec = d_vec.download(np.float32, dim)
                key = (int(info.iter_count), int(np.float32(info.eigen_val).view(np.uint32)), hash(vec.tobytes()))
                if cap not in first:
                    first[cap] = key
                assert first[cap] == key, (kind, dim, cap, solves, first[cap], key)
                solves += 1
            if rank == 0 and time.time() - t0 > budget:
                stop.fill_(1)
            dist.broadcast(stop, 0)
            if int(stop.item()):
                break
        # all ranks hold the same results
        mine = torch.tensor([k[1] for k in (first[c] for c in caps)], dtype=torch.int64, device="cuda")
        ref = mine.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(mine, ref), (kind, dim)
        report.append({"case": f"{kind}-{dim}", "solves": solves, "iter_count_by_cap": {str(c): first[c][0] for c in caps}})
        d_rows.free()
        d_vec.free()
        sh.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps({"tool": "stress_sharded", "world": world, "seconds": args.seconds, "cases": report,
                          "result": "every solve bit-identical to the first of its kind on every rank"}), flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
