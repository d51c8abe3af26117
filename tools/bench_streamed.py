#!/usr/bin/env python
"""Streamed solve (st_solve_streamed) against the in-device solve of the same matrix, one JSON line.

    python tools/bench_streamed.py [--dim 32768] [--cached 0.5] [--block-rows 0] [--pinned 1] [--steps 3]

The matrix is a host Hilbert matrix (pinned with st_pin_host unless --pinned 0).  `--cached` is the share
of the matrix the device cache may hold (the rest crosses PCIe every round), so one GPU can play "matrix
larger than HBM" at a size that finishes in seconds.  Reported: ms per round, GB/s of matrix consumed
per round (4*N^2 bytes per pass), PCIe GB/s (bytes actually copied / time), and the expected floor
max(cached bytes / HBM rate, streamed bytes / PCIe rate) from MEASURED_PEAKS.json and 55 GB/s.
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
    ap.add_argument("--dim", type=int, default=32768)
    ap.add_argument("--cached", type=float, default=0.5)
    ap.add_argument("--block-rows", type=int, default=0)
    ap.add_argument("--pinned", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()

    import numpy as np
    from eigen_value_b200 import Solver
    from eigen_value_b200._lib import check

    solver = Solver(0)
    dim = args.dim
    d = solver.hilbert(dim)
    solver.synchronize()
    mat = np.empty((dim, dim), dtype=np.float32)
    check(solver.lib.st_memcpy_d2h(solver.ctx, mat.ctypes.data, d.ptr, mat.nbytes), "d2h")
    base, base_vec = solver.solve_device(d, dim)
    d.free()
    if args.pinned:
        check(solver.lib.st_pin_host(solver.ctx, mat.ctypes.data, mat.nbytes), "st_pin_host")
    budget = max(int(args.cached * mat.nbytes), 2 * 4 * dim)
    runs = []
    for _ in range(1 + args.steps):
        t0 = time.perf_counter()
        info, vec, plan = solver.solve_streamed(mat, device_budget=budget, block_rows=args.block_rows)
        runs.append(((time.perf_counter() - t0) * 1e3, info, plan))
    if args.pinned:
        check(solver.lib.st_unpin_host(solver.ctx, mat.ctypes.data), "st_unpin_host")
    same = bool(info.iter_count == base.iter_count and np.float32(info.eigen_val) == np.float32(base.eigen_val)
                and np.array_equal(vec, base_vec))
    wall_ms, info, plan = min(runs[1:], key=lambda r: r[1].loop_ms)
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    pcie = 55.0
    per_round = plan["h2d_bytes_per_round"]
    cached_bytes = 4 * dim * dim - per_round
    floor_ms = max(cached_bytes / (hbm * 1e9), per_round / (pcie * 1e9)) * 1e3
    later_ms = (info.loop_ms - 0.0) / info.passes
    print(json.dumps({
        "tool": "bench_streamed", "workload": f"hilbert-{dim}", "pinned": bool(args.pinned), "plan": plan,
        "rounds": info.iter_count, "passes": info.passes, "loop_ms": round(info.loop_ms, 3),
        "ms_per_round_mean": round(later_ms, 3), "us_per_round_median": round(info.round_us_median, 1),
        "matrix_gbs_per_round": round(4.0 * dim * dim / (info.round_us_median * 1e-6) / 1e9, 1),
        "pcie_gbs": round(plan["h2d_bytes_total"] / (info.loop_ms * 1e-3) / 1e9, 2),
        "floor_ms_per_round": round(floor_ms, 3), "floor_assumes": {"hbm_gbs": hbm, "pcie_gbs": pcie},
        "in_device_loop_ms": round(base.loop_ms, 3), "bit_identical_to_in_device_solve": same,
        "launches": info.launches, "wall_ms": round(wall_ms, 1)}))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
