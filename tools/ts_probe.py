"""Per-phase view of small solves: python tools/ts_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigen_value_b200 import Solver  # noqa: E402

s = Solver(0)
for n in (512, 1024, 1536, 2048, 2304, 4096):
    d = s.hilbert(n)
    for kernel in (0, 13, 1):
        best = None
        for _ in range(6):
            info, _ = s.solve_device(d, n, kernel=kernel)
            if best is None or info.loop_ms < best[0].loop_ms:
                best = (info, s.phase_breakdown())
        info, ph = best
        print(n, "kernel", info.kernel_id, "grid", info.grid, "rounds", info.iter_count,
              "loop_us %.1f" % (info.loop_ms * 1e3), "round_med %.2f" % info.round_us_median,
              {k: round(v, 2) for k, v in ph.items()})
