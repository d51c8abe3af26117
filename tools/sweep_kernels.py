"""Kernel-variant sweep over the st_options.kernel ids that exist (round 2 removed the TMA-ring, 256/1024-thread and L2-prefetch variants).
Usage: python tools/sweep_kernels.py [N ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eigen_value_b200 import Solver  # noqa: E402

PEAK = 6554.9
NAMES = {0: "auto", 1: "general ldg", 10: "sc 512 pf2", 11: "sc 512 pf0", 12: "sc 512 pf3", 13: "sc 512 pf1", 20: "cluster"}
SWEEPS = [int(x) for x in os.environ.get("SWEEPS", "1,0").split(",") if x]   # st_options.sweep values (3 = alternating + static)
ONLY = [int(x) for x in os.environ.get("KERNELS", "").split(",") if x]


def main():
    dims = [int(a) for a in sys.argv[1:]] or [8192]
    s = Solver(0)
    print(f"# {s.name} SMs={s.sm_count} L2={s.l2_bytes / 2**20:.0f} MiB")
    for dim in dims:
        d = s.hilbert(dim)
        ref = None
        for sweep in SWEEPS:
            for kid in (ONLY or sorted(NAMES)):
                try:
                    best = None
                    for rep in range(4):
                        info, vec = s.solve_device(d, dim, sweep=sweep, kernel=kid, max_iter=40)
                        if best is None or info.loop_ms < best.loop_ms:
                            best = info
                except Exception as exc:
                    print(f"N={dim} sweep={sweep} kernel={kid} ({NAMES[kid]}): {exc}")
                    continue
                if ref is None:
                    ref = (best.eigen_val, vec)
                same = best.eigen_val == ref[0] and (vec == ref[1]).all()
                gbs = best.bytes_per_round / (best.round_us_median * 1e-6) / 1e9
                tot = best.bytes_per_round * best.passes / (best.loop_ms * 1e-3) / 1e9
                print(f"N={dim} sweep={sweep} kernel={kid} ({NAMES[kid]:16s}) rounds={best.iter_count} "
                      f"loop={best.loop_ms * 1e3:9.1f}us round_med={best.round_us_median:8.2f}us "
                      f"min={best.round_us_min:8.2f}us GB/s(loop)={tot:6.0f} ({tot / PEAK:.3f}) "
                      f"bitwise_same={same}", flush=True)
        d.free()


if __name__ == "__main__":
    main()
