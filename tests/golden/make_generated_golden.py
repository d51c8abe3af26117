#!/usr/bin/env python
"""Expected results for the BASELINE sizes no host can hold as a matrix, computed on the CPU by the oracle's
matrix-free loop (oracle_similarity_transform_generated: rows are generated and reduced one at a time, in the CUDA
kernels' evaluation order, ORACLE_SUM_CUDA).  Writes tests/golden/generated_expected.json; bench.py compares the
GPU's eigenvalue bits, round count and eigenvector digest against it (its `parity` record), and
tests/test_gpu_parity.py does the same for the sizes that fit one GPU.

    python tests/golden/make_generated_golden.py [--threads 6] [--only hilbert-65536]

Minutes per case on 8 host cores: hilbert-65536 ~1, hilbert-131072 ~3, uniform capped at 50 rounds ~5 / ~20.
"""
import argparse
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "generated_expected.json")
CASES = [  # (workload, kind, dim, seed, max_itr)
    ("hilbert-16384", "hilbert", 16384, 0, 1000),
    ("hilbert-32768", "hilbert", 32768, 0, 1000),
    ("hilbert-65536", "hilbert", 65536, 0, 1000),
    ("hilbert-131072", "hilbert", 131072, 0, 1000),
    ("uniform-32768", "uniform", 32768, 0x5EED0000 + 32768, 50),
    ("uniform-65536", "uniform", 65536, 0x5EED0001, 50),
    ("uniform-131072", "uniform", 131072, 0x5EED0002, 50),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if args.threads:
        oracle.lib().oracle_set_threads(args.threads)
    doc = {"about": "CPU-computed expected results (oracle.similarity_transform_generated, ORACLE_SUM_CUDA order, read-only form, "
                    "eps 1e-3, absolute stop) for generated matrices; eigen_val is the float32 widened to double, "
                    "eigen_vec_sha256 the digest of the raw float32 eigenvector bytes. Made by tests/golden/make_generated_golden.py.",
           "cases": {}}
    if os.path.exists(OUT):
        with open(OUT) as f:
            doc = json.load(f)
    for name, kind, dim, seed, max_itr in CASES:
        if args.only and name != args.only:
            continue
        if name in doc["cases"] and not args.only:
            continue
        t0 = time.time()
        val, vec, _, it = oracle.similarity_transform_generated(kind, dim, seed, max_itr=max_itr)
        doc["cases"][name] = {
            "kind": kind, "dim": dim, "seed": seed, "max_iter": max_itr, "iter_count": int(it),
            "eigen_val": float(val), "eigen_val_bits": int(np.float32(val).view(np.uint32)),
            "eigen_vec_sha256": hashlib.sha256(np.ascontiguousarray(vec).tobytes()).hexdigest(),
            "eigen_vec_head": [float(x) for x in vec[:4]], "eigen_vec_max": float(vec.max()),
            "cpu_seconds": round(time.time() - t0, 1), "threads": oracle.threads()}
        with open(OUT, "w") as f:
            json.dump(doc, f, indent=1)
        print(name, doc["cases"][name], flush=True)


if __name__ == "__main__":
    main()
