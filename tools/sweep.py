"""Tuning sweep on one GPU: per-round time of the round-loop kernel for several launch shapes.
Usage: python tools/sweep.py [N ...]   (writes a table to stdout)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eigen_value_b200 import Solver  # noqa: E402

PEAK = 6554.9


def main():
    dims = [int(a) for a in sys.argv[1:]] or [8192]
    s = Solver(0)
    print(f"# {s.name} SMs={s.sm_count} L2={s.l2_bytes/2**20:.0f} MiB")
    for dim in dims:
        d = s.hilbert(dim)
        for form in (0, 1):
            for threads in (256, 512, 1024):
                for sweep in (0, 1):
                    best = None
                    for rep in range(4):
                        info, _ = s.solve_device(d, dim, form=form, threads=threads, sweep=sweep)
                        if best is None or info.loop_ms < best.loop_ms:
                            best = info
                    gbs = best.bytes_per_round / (best.round_us_median * 1e-6) / 1e9
                    tot = best.bytes_per_round * best.passes / (best.loop_ms * 1e-3) / 1e9
                    print(f"N={dim} form={form} threads={threads} sweep={sweep} grid={best.grid} "
                          f"rounds={best.iter_count} loop={best.loop_ms*1e3:.1f}us "
                          f"round_med={best.round_us_median:.2f}us round_min={best.round_us_min:.2f}us "
                          f"GB/s(med round)={gbs:.0f} ({gbs/PEAK:.2f}) GB/s(loop)={tot:.0f} ({tot/PEAK:.2f}) "
                          f"lambda={best.eigen_val:.7f}", flush=True)
        d.free()


if __name__ == "__main__":
    main()
