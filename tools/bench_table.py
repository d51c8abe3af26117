"""The reference's benchmark table (main.cpp:23-35 -> README.md:70-76), produced by this library:
Hilbert N x N for N = 2^7 ... 2^13 (and beyond with --max-log2), one line per size in the
reference's own format, through the drop-in entry point (host matrix in, like
benchmarks/benchmark_similarity_transform.cpp:3-22) and, next to it, device-resident.

    python tools/bench_table.py [--max-log2 15]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from eigen_value_b200 import EigenValue, Solver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=13)
    args = ap.parse_args()
    ev = EigenValue()
    solver = Solver(0)
    print(f"running on {solver.name}\n")
    print("Parallel Similarity Transform for finding max eigen value (with vector)\n")
    for i in range(7, args.max_log2 + 1):
        dim = 1 << i
        d = solver.hilbert(dim)                       # generate_hilbert_matrix on the device
        best = None
        for _ in range(3):
            info, _ = solver.solve_device(d, dim)
            best = info if best is None or info.loop_ms < best.loop_ms else best
        line = f"{dim:<5}x{dim:>5}\t\t\t{best.loop_ms:10.3f} ms\t\t\t{best.iter_count:>6} round(s)"
        if dim <= 8192:                               # the reference's path: host matrix through max_eigen_value
            host = d.download(np.float32, dim * dim).reshape(dim, dim)
            lam, vec, ms, itr = ev.similarity_transform(host)
            assert itr == best.iter_count
            line += f"\t\tmax_eigen_value(): {ms:>4} ms (loop, whole ms), lambda = {lam:.7f}"
        print(line)
        d.free()


if __name__ == "__main__":
    main()
