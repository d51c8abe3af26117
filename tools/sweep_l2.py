"""L2-residency experiment: per-round time vs share of rows loaded evict_last, with and
without the alternating sweep.  Usage: python tools/sweep_l2.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eigen_value_b200 import Solver  # noqa: E402

PEAK = 6554.9


def main():
    dims = [int(a) for a in sys.argv[1:]] or [8192]
    s = Solver(0)
    print(f"# {s.name} SMs={s.sm_count} L2={s.l2_bytes / 2**20:.0f} MiB")
    for dim in dims:
        d = s.hilbert(dim)
        ref = None
        for sweep in (0, 1):
            for keep in (0, 15, 25, 30, 35, 40, 45, 50, 60):
                best = None
                for rep in range(4):
                    info, vec = s.solve_device(d, dim, sweep=sweep, l2_keep_pct=keep, max_iter=40)
                    if best is None or info.loop_ms < best.loop_ms:
                        best = info
                if ref is None:
                    ref = (best.eigen_val, vec)
                same = best.eigen_val == ref[0] and (vec == ref[1]).all()
                gbs = best.bytes_per_round / (best.round_us_median * 1e-6) / 1e9
                tot = best.bytes_per_round * best.passes / (best.loop_ms * 1e-3) / 1e9
                print(f"N={dim} sweep={sweep} keep={keep:3d}% rounds={best.iter_count} "
                      f"loop={best.loop_ms * 1e3:9.1f}us round_med={best.round_us_median:8.2f}us "
                      f"min={best.round_us_min:8.2f}us GB/s(med)={gbs:6.0f} ({gbs / PEAK:.2f}) "
                      f"GB/s(loop)={tot:6.0f} ({tot / PEAK:.2f}) bitwise_same={same}", flush=True)
        d.free()


if __name__ == "__main__":
    main()
