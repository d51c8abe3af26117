#!/usr/bin/env python
"""End-to-end max_eigen_value() through the drop-in boundary, one GPU vs a device group, one JSON line.

    gpurun --gpus 8 -- python tools/bench_group.py [--dim 8192] [--steps 5] [--pinned 1]

The same host matrix (Hilbert) goes through make_queue / max_eigen_value exactly as the reference's wrapper
calls them: first with the handle alone, then with every other GPU attached (st_group_attach; what
ST_DEVICES=all does for the unmodified wrapper).  Host->device copies are inside the timed region, like
bench.py's e2e leg.  With a group every GPU's PCIe link carries 1/G of the matrix, so the copy that
dominates the one-GPU call shrinks with G.  Results must be bit-identical.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--pinned", type=int, default=1)
    ap.add_argument("--min-dim", type=int, default=0)
    args = ap.parse_args()

    import numpy as np
    from eigen_value_b200 import EigenValue, Solver
    from eigen_value_b200._lib import check

    dim = args.dim
    gen = Solver(0)
    d = gen.hilbert(dim)
    gen.synchronize()
    mat = np.empty((dim, dim), dtype=np.float32)
    check(gen.lib.st_memcpy_d2h(gen.ctx, mat.ctypes.data, d.ptr, mat.nbytes), "d2h")
    d.free()
    if args.pinned:
        check(gen.lib.st_pin_host(gen.ctx, mat.ctypes.data, mat.nbytes), "st_pin_host")

    def timed(ev):
        best, out = None, None
        for i in range(1 + args.steps):
            t0 = time.perf_counter()
            out = ev.similarity_transform(mat)
            ms = (time.perf_counter() - t0) * 1e3
            if i and (best is None or ms < best):
                best = ms
        return best, out

    one = EigenValue()
    one_ms, (v1, e1, loop1, it1) = timed(one)
    grp = EigenValue(devices="all", min_dim=args.min_dim)
    n = grp.device_count
    grp_ms, (v2, e2, loop2, it2) = timed(grp)
    passes = min(it1 + 1, 1000)
    gbs = lambda ms: round(passes * 4.0 * dim * dim / (ms * 1e-3) / 1e9, 1)
    same = bool(it1 == it2 and v1 == v2 and np.array_equal(e1, e2))
    print(json.dumps({
        "tool": "bench_group", "workload": f"hilbert-{dim}", "pinned": bool(args.pinned), "rounds": it1,
        "one_gpu": {"e2e_ms": round(one_ms, 3), "e2e_gbs": gbs(one_ms), "loop_ms": int(loop1)},
        "group": {"gpus": n, "e2e_ms": round(grp_ms, 3), "e2e_gbs": gbs(grp_ms), "loop_ms": int(loop2)},
        "speedup": round(one_ms / grp_ms, 2), "h2d_bytes_per_gpu": 4 * dim * dim // max(n, 1),
        "bit_identical": same}))
    if args.pinned:
        check(gen.lib.st_unpin_host(gen.ctx, mat.ctypes.data), "st_unpin_host")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
