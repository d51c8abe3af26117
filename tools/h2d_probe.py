#!/usr/bin/env python
"""Host -> device copy rate of this box with 1, 2, 4, ... GPUs copying at once (pinned 1 GiB per GPU, one host thread
per GPU, st_memcpy_h2d): the ceiling of every end-to-end number that starts from a host matrix.  One JSON line.

    gpurun --gpus 8 -- python tools/h2d_probe.py [--mib 1024] [--reps 3]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from eigen_value_b200 import Solver
    from eigen_value_b200._lib import check

    n = min(torch.cuda.device_count(), 8)
    nbytes = args.mib << 20
    solvers = [Solver(g) for g in range(n)]
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
    dev = [s.alloc(nbytes) for s in solvers]
    out = {"tool": "h2d_probe", "mib_per_gpu": args.mib, "cpus": len(os.sched_getaffinity(0)), "runs": []}
    g = 1
    while g <= n:
        best = None
        for _ in range(args.reps):
            start = threading.Barrier(g + 1)
            done = [0.0] * g

            def work(i):
                start.wait()
                t0 = time.perf_counter()
                check(solvers[i].lib.st_memcpy_h2d(solvers[i].ctx, dev[i].ptr, host[i].data_ptr(), nbytes), "h2d")
                done[i] = time.perf_counter() - t0

            th = [threading.Thread(target=work, args=(i,)) for i in range(g)]
            for t in th:
                t.start()
            start.wait()
            t0 = time.perf_counter()
            for t in th:
                t.join()
            wall = time.perf_counter() - t0
            if best is None or wall < best[0]:
                best = (wall, list(done))
        out["runs"].append({"gpus": g, "wall_ms": round(best[0] * 1e3, 2),
                            "aggregate_gbs": round(g * nbytes / best[0] / 1e9, 1),
                            "per_gpu_gbs": [round(nbytes / d / 1e9, 1) for d in best[1]]})
        g *= 2
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
