"""Per-kernel benchmark tables in the reference's format (main.cpp:37-159, producing
benchmarks/similarity_transform.md), for the B200 build.

The reference times each L1 kernel alone: `sum_across_rows` and `compute_next_matrix` on N x N
matrices for N = 2^7..2^13 (main.cpp:39,138), the three vector kernels on vectors of 2^16..2^25
floats (main.cpp:78,114,151).  Here the same entry points are the `st_*` per-kernel calls of the
C ABI, timed on the device with the CUDA-event stopwatch (st_timer_start / st_timer_stop): one
warm-up call, then the mean of `--reps` back-to-back calls.  The fused round loop does not run these
kernels (it makes one pass per round); the last table therefore splits a fused
round into its phases from the in-kernel stamps (st_phase_timestamps), which is the closest
counterpart the fused design has to a per-kernel figure.

    python tools/bench_kernels.py [--reps 20] [--max-log2 13] [--max-vec-log2 25] [--json out.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from eigen_value_b200 import Solver  # noqa: E402
from eigen_value_b200._lib import check  # noqa: E402


def timed(solver, reps, call):
    call()                                   # warm-up (module load, caches)
    solver.synchronize()
    solver.timer_start()
    for _ in range(reps):
        call()
    return solver.timer_stop() / reps        # ms per call, device time


def row(dim, ms, extra=""):
    print(f"{dim:<5}x{dim:>5}\t\t\t{ms:10.4f} ms{extra}")


def vrow(n, ms, extra=""):
    print(f"{n:>10}\t\t\t{ms:10.4f} ms{extra}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--max-log2", type=int, default=13)
    ap.add_argument("--max-vec-log2", type=int, default=25)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    s = Solver(0)
    lib, ctx = s.lib, s.ctx
    out = {"device": s.name, "reps": args.reps, "tables": {}}
    print(f"running on {s.name}\n")

    mats = range(7, args.max_log2 + 1)
    vecs = range(16, args.max_vec_log2 + 1)

    # ---- sum_across_rows, reference main.cpp:37-74 (v0/v1/v2 are one kernel here) ----
    print("[kernel] Sum Across Rows of Matrix\n")
    tab = []
    for i in mats:
        dim = 1 << i
        d_mat, d_vec = s.uniform(dim, 0x5EED0000 + dim), s.alloc(4 * dim)
        ms = timed(s, args.reps, lambda: check(lib.st_sum_across_rows(ctx, d_mat.ptr, d_vec.ptr, dim), "sum"))
        gbs = 4.0 * dim * dim / (ms * 1e-3) / 1e9
        row(dim, ms, f"\t\t{gbs:8.1f} GB/s")
        tab.append({"N": dim, "ms": ms, "GBps": gbs})
        d_mat.free(), d_vec.free()
    out["tables"]["sum_across_rows"] = tab

    # ---- find_max, reference main.cpp:76-110 ----
    print("\n[kernel] Max Value in Vector\n")
    tab = []
    for i in vecs:
        n = 1 << i
        d_vec, d_max = s.upload(np.arange(1, n + 1, dtype=np.float32)), s.alloc(4)
        ms = timed(s, args.reps, lambda: check(lib.st_find_max(ctx, d_vec.ptr, d_max.ptr, n), "max"))
        assert d_max.download(np.float32, 1)[0] == np.float32(n)
        vrow(n, ms)
        tab.append({"n": n, "ms": ms})
        d_vec.free(), d_max.free()
    out["tables"]["find_max"] = tab

    # ---- compute_eigen_vector, reference main.cpp:112-134 ----
    print("\n[kernel] Eigen Vector Computation\n")
    tab = []
    for i in vecs:
        n = 1 << i
        d_vec, d_max = s.upload(np.ones(n, dtype=np.float32)), s.upload(np.array([1.0], dtype=np.float32))
        d_e = s.upload(np.ones(n, dtype=np.float32))
        ms = timed(s, args.reps, lambda: check(lib.st_compute_eigen_vector(ctx, d_vec.ptr, d_max.ptr, d_e.ptr, n), "ev"))
        vrow(n, ms)
        tab.append({"n": n, "ms": ms})
        d_vec.free(), d_max.free(), d_e.free()
    out["tables"]["compute_eigen_vector"] = tab

    # ---- compute_next_matrix, reference main.cpp:136-147 ----
    print("\n[kernel] Next Matrix Computation\n")
    tab = []
    for i in mats:
        dim = 1 << i
        d_mat, d_vec = s.uniform(dim, 0x5EED0000 + dim), s.upload(np.ones(dim, dtype=np.float32))
        ms = timed(s, args.reps, lambda: check(lib.st_compute_next_matrix(ctx, d_mat.ptr, d_vec.ptr, dim), "next"))
        gbs = 8.0 * dim * dim / (ms * 1e-3) / 1e9
        row(dim, ms, f"\t\t{gbs:8.1f} GB/s (read + write)")
        tab.append({"N": dim, "ms": ms, "GBps": gbs})
        d_mat.free(), d_vec.free()
    out["tables"]["compute_next_matrix"] = tab

    # ---- stop, reference main.cpp:149-159 ----
    print("\n[kernel] Stop Criteria Checker\n")
    tab = []
    for i in vecs:
        n = 1 << i
        d_vec, d_ret = s.upload(np.ones(n, dtype=np.float32)), s.alloc(4)
        ms = timed(s, args.reps, lambda: check(lib.st_stop(ctx, d_vec.ptr, d_ret.ptr, n, 1e-3), "stop"))
        assert d_ret.download(np.uint32, 1)[0] == 1
        vrow(n, ms)
        tab.append({"n": n, "ms": ms})
        d_vec.free(), d_ret.free()
    out["tables"]["stop"] = tab

    # ---- the fused round, split into phases (no counterpart in the reference: it has no fused loop) ----
    print("\n[fused round loop] per-round phases of similarity_transform(), Hilbert N x N, microseconds\n")
    print(f"{'N':>6} {'rounds':>7} {'pass':>9} {'barrier':>9} {'tail':>9} {'round':>9}")
    tab = []
    for i in mats:
        dim = 1 << i
        d = s.hilbert(dim)
        best = None
        for _ in range(3):
            info, _ = s.solve_device(d, dim)
            ph = s.phase_breakdown()
            if best is None or info.loop_ms < best[0].loop_ms:
                best = (info, ph)
        info, ph = best
        print(f"{dim:>6} {info.iter_count:>7} {ph['pass_us']:>9.2f} {ph['barrier_us']:>9.2f} {ph['tail_us']:>9.2f} "
              f"{info.round_us_median:>9.2f}")
        tab.append({"N": dim, "rounds": info.iter_count, **ph, "round_us": info.round_us_median,
                    "kernel": info.kernel_name, "kernel_id": info.kernel_id})
        d.free()
    out["tables"]["fused_round_phases"] = tab

    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
