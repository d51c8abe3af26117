"""Generates tests/golden/reference_sycl.json from the UNMODIFIED reference C++ sources running on
the CPU SYCL shim (oracle/_ref/libreference_cpu.so; build it with `make -C oracle ref`, which
needs /root/reference and therefore only works in the build container).

    python tests/golden/make_reference_golden.py

Every case records the input recipe (so the matrix can be rebuilt anywhere, bit for bit), the
work-group size the reference's wrapper picked (wrapper/similarity_transform.cpp:33), and the
reference's outputs: lambda, the raw eigenvector (base64 of the float32 bytes) and iter_count.
"""
import base64
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import ref  # noqa: E402


def build_matrix(recipe):
    kind = recipe["kind"]
    if kind == "literal":
        return np.array(recipe["rows"], dtype=np.float32)
    if kind == "hilbert":
        return oracle.hilbert(recipe["dim"])
    if kind == "uniform":   # Philox uniform (0,1] + offset, oracle.c:oracle_generate_uniform
        return (oracle.uniform(recipe["dim"], recipe["seed"]) + np.float32(recipe["offset"])).astype(np.float32)
    raise ValueError(kind)


def main():
    if not ref.available():
        raise SystemExit("oracle/_ref/libreference_cpu.so missing: run `make -C oracle ref` first")
    recipes = [
        {"name": "golden3x3_wg3", "kind": "literal", "rows": [[1, 1, 2], [2, 1, 3], [2, 3, 5]], "wg": 3},
        {"name": "golden3x3_wrapper", "kind": "literal", "rows": [[1, 1, 2], [2, 1, 3], [2, 3, 5]]},
    ]
    for n in (128, 256, 512, 1024):
        recipes.append({"name": f"hilbert{n}", "kind": "hilbert", "dim": n})
    for n, off in ((8, 0.05), (64, 0.05), (100, 0.25), (250, 0.25), (512, 0.0), (1024, 0.0)):
        recipes.append({"name": f"uniform{n}", "kind": "uniform", "dim": n, "seed": 0x5EED0000 + n, "offset": off})
    cases = []
    for rc in recipes:
        m = build_matrix(rc)
        n = m.shape[0]
        if "wg" in rc:
            wg = rc["wg"]
            val, vec, _, it = ref.similarity_transform(m, wg)      # tests/test.cpp:96-97 path
        else:
            wg = ref.wrapper_wg_size(n)
            val, vec, _, it = ref.max_eigen_value(m)               # Python wrapper path
        case = dict(rc)
        case.update({"wg_size": wg, "eigen_val": float(val), "eigen_val_hex": float(val).hex(),
                     "iter_count": it, "eigen_vec_b64": base64.b64encode(vec.tobytes()).decode(),
                     "matrix_checksum": float(np.float64(m.astype(np.float64).sum()))})
        cases.append(case)
        print(rc["name"], "wg", wg, "lambda", val, "iters", it)
    # a launch shape the reference itself rejects (dim % wg_size != 0), SURVEY 8(b)
    rejected = []
    for n in (5, 640, 1000):
        try:
            ref.max_eigen_value(oracle.uniform(n, 1) + np.float32(0.1))
            rejected.append({"dim": n, "rejected": False})
        except RuntimeError:
            rejected.append({"dim": n, "rejected": True, "wg_size": ref.wrapper_wg_size(n)})
    out = {"generator": "tests/golden/make_reference_golden.py",
           "reference": "similarity_transform.cpp + wrapper/similarity_transform.cpp (unmodified) on oracle/sycl_shim",
           "sub_group_size": 32, "max_work_group_size": int(ref.lib().ref_max_work_group_size()),
           "cases": cases, "rejected_shapes": rejected}
    with open(os.path.join(HERE, "reference_sycl.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
