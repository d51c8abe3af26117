"""Helpers shared by the golden-fixture tests (tests/golden/reference_sycl.json)."""
import base64
import json
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_reference_cases():
    with open(os.path.join(GOLDEN, "reference_sycl.json")) as f:
        doc = json.load(f)
    return doc


def build_matrix(case) -> np.ndarray:
    kind = case["kind"]
    if kind == "literal":
        m = np.array(case["rows"], dtype=np.float32)
    elif kind == "hilbert":
        m = oracle.hilbert(case["dim"])
    elif kind == "uniform":
        m = (oracle.uniform(case["dim"], case["seed"]) + np.float32(case["offset"])).astype(np.float32)
    else:
        raise ValueError(kind)
    # the recipe rebuilt the very matrix the reference was run on
    assert float(np.float64(m.astype(np.float64).sum())) == case["matrix_checksum"]
    return m


def expected(case):
    vec = np.frombuffer(base64.b64decode(case["eigen_vec_b64"]), dtype=np.float32)
    return np.float32(float.fromhex(case["eigen_val_hex"])), vec, case["iter_count"]
