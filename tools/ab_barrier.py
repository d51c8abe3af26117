#!/usr/bin/env python
"""A/B of the two cross-GPU round barriers in ONE process group, same build, same box: the flat barrier (every CTA
arrives at every GPU's counter itself; st_options.sweep = 1, or 33 with the sequentially consistent fence) against the
forwarding-flag protocol (sweep = 17).  Alternates them, several laps per size; loop time = max over ranks of the CUDA-event time of the launch.
One JSON line from rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 \
        tools/ab_barrier.py [--dims 32768,16384] [--solves 10] [--laps 3]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from eigen_value_b200 import Solver  # noqa: E402
from eigen_value_b200.sharded import ShardedSolver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="32768,16384")
    ap.add_argument("--solves", type=int, default=10)
    ap.add_argument("--laps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solver = Solver(local)
    out = []
    for dim in (int(x) for x in args.dims.split(",")):
        sh = ShardedSolver(solver, dim, rank, world)
        d_rows = sh.hilbert()
        d_vec = solver.alloc(4 * dim)
        variants = {1: "flat, release/acquire fence", 33: "flat, sequentially consistent fence", 17: "forwarding flags"}
        for sweep in variants:                      # warm-up of each
            sh.solve(d_rows, d_eigen_vec=d_vec, sweep=sweep)
        bits = None
        for lap in range(args.laps):
            for sweep in variants:
                us, phases = [], []
                for _ in range(args.solves):
                    dist.barrier()
                    info, _ = sh.solve(d_rows, d_eigen_vec=d_vec, sweep=sweep)
                    t = torch.tensor([info.loop_ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    us.append(float(t.item()) * 1e3 / info.passes)
                    phases.append(solver.phase_breakdown())
                    b = (int(np.float32(info.eigen_val).view(np.uint32)), int(info.iter_count))
                    bits = bits or b
                    assert b == bits, (dim, sweep, b, bits)
                med = lambda k: float(np.median([p[k] for p in phases if p]))  # noqa: E731
                out.append({"dim": dim, "barrier": variants[sweep], "lap": lap,
                            "us_per_round_median": round(float(np.median(us)), 3), "us_per_round_min": round(min(us), 3),
                            "rank0_phase_us": {k: round(med(k), 3) for k in ("pass_us", "barrier_us", "tail_us")},
                            "rounds": int(info.iter_count), "eigen_val_bits": bits[0]})
        d_rows.free()
        d_vec.free()
        sh.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps({"tool": "ab_barrier", "world": world, "solves_per_lap": args.solves, "records": out}), flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
