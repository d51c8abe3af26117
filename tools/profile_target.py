"""Smallest program that runs the dominant kernel the way bench.py does: N solves of one
workload on cuda:0.  Used as the ncu target (profiles/README.md has the command lines).

    python tools/profile_target.py [workload] [solves] [key=value ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eigen_value_b200 import Solver  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "hilbert-8192"
    solves = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    opts = {"sweep": 1}
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        opts[k] = float(v) if k == "eps" else int(v)
    kind, n = workload.split("-")
    n = int(n)
    s = Solver(0)
    d = s.hilbert(n) if kind == "hilbert" else s.uniform(n, 0x5EED0000 + n)
    v = s.alloc(4 * n)
    for i in range(solves):
        info, _ = s.solve_device(d, n, d_eigen_vec=v, **opts)
        gbs = info.bytes_per_round * info.passes / (info.loop_ms * 1e-3) / 1e9
        print(f"{workload} solve {i}: rounds={info.iter_count} loop={info.loop_ms * 1e3:.1f} us "
              f"round_med={info.round_us_median:.2f} us  {gbs:.0f} GB/s  lambda={info.eigen_val:.7f}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
