"""Generates tests/golden/main_py.json by importing the UNMODIFIED reference main.py from
/root/reference (only possible in the build container; the GPU box has no reference tree).

    python tests/golden/make_golden.py

main.py is the reference's sequential fp64 model (main.py:30-47).  Its stop rule is
non-circular and it counts itr+1, so its round counts are NOT comparable with the SYCL path
(SURVEY 0.3); the fixture pins the eigenpair it converges to, on small positive matrices.
"""
import importlib.util
import json
import os
import sys

import numpy as np

REFERENCE = os.environ.get("REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_main():
    spec = importlib.util.spec_from_file_location("reference_main", os.path.join(REFERENCE, "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)   # the __main__ block does not run on import
    return mod


def main():
    ref = load_reference_main()
    cases = []
    mats = {"golden3x3": np.array([[1, 1, 2], [2, 1, 3], [2, 3, 5]], dtype=np.float32)}
    rng = np.random.default_rng(20211018)
    for n in (8, 32, 64):
        mats[f"random{n}"] = rng.random((n, n)).astype(np.float32) + np.float32(0.01)
    r, c = np.indices((48, 48))
    mats["hilbert48"] = (np.float32(1.0) / (r + c + 1).astype(np.float32)).astype(np.float32)
    for name, m in mats.items():
        val, vec, rounds = ref.max_eigen_value_and_vector(m)      # fp64 inside (numpy upcasts)
        cases.append({"name": name, "matrix": m.tolist(), "eigen_val": float(val),
                      "eigen_vec": [float(x) for x in vec], "rounds_main_py": int(rounds)})
    out = {"generator": "tests/golden/make_golden.py", "reference": "main.py:max_eigen_value_and_vector",
           "numpy": np.__version__, "cases": cases}
    with open(os.path.join(HERE, "main_py.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    sys.exit(main())
