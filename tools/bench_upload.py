#!/usr/bin/env python
"""max_eigen_value() end to end on a PAGEABLE host matrix (what the reference's wrapper passes: a numpy array),
with the driver's staging vs the opt-in multi-threaded upload, one JSON line per setting.

    python tools/bench_upload.py [--dim 8192] [--steps 5] [--threads 0,2,4,8]

Each setting runs in a fresh process because ST_UPLOAD_THREADS is read when the handle is created; the pinned
figure (st_pin_host) is printed beside them as the ceiling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(dim: int, steps: int) -> int:
    import oracle
    from eigen_value_b200 import EigenValue

    mat = oracle.hilbert(dim)                       # plain numpy allocation: pageable
    ev = EigenValue()

    def best(fn):
        out = []
        for i in range(1 + steps):
            t0 = time.perf_counter()
            r = fn()
            if i:
                out.append((time.perf_counter() - t0) * 1e3)
        return min(out), r

    ms, (val, vec, loop_ms, it) = best(lambda: ev.similarity_transform(mat))
    with ev.pinned(mat):
        pinned_ms, _ = best(lambda: ev.similarity_transform(mat))
    print(json.dumps({"tool": "bench_upload", "workload": f"hilbert-{dim}",
                      "upload_threads": int(os.environ.get("ST_UPLOAD_THREADS", "0") or 0),
                      "pageable_e2e_ms": round(ms, 3), "pinned_e2e_ms": round(pinned_ms, 3),
                      "pageable_gbs_h2d": round(mat.nbytes / (ms * 1e-3) / 1e9, 2),
                      "staged_bytes": int(ev.so_lib.st_staged_upload_bytes(ev.sycl_q)),
                      "rounds": it, "lambda": float(val)}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", default="0,2,4,8")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args.dim, args.steps)
    rc = 0
    for t in args.threads.split(","):
        env = dict(os.environ, ST_UPLOAD_THREADS=t.strip())
        proc = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--dim", str(args.dim),
                               "--steps", str(args.steps)], env=env, cwd=ROOT)
        rc = rc or proc.returncode
    return rc


if __name__ == "__main__":
    sys.exit(main())
