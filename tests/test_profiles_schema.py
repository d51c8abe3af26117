"""The bench lines committed under profiles/ carry every key of the measurement contract (so a
change of bench.py that drops one is caught on the CPU), and their numbers are self-consistent."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_bench_n1_hilbert8192.json"))
               + glob.glob(os.path.join(ROOT, "profiles", "r1_scale_*.json"))
               + glob.glob(os.path.join(ROOT, "profiles", "r2_bench_n[1248].json")))      # the final build, driver-style lines

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e",
             "gpu_launches", "clocks"}


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_bench_line_has_the_contract_keys(path):
    with open(path) as f:
        d = json.loads(f.read())
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert abs(r["achieved"] - d["value"]) < 1e-6 * d["value"]
    # value = passes * 4 N^2 bytes / device time, whole job
    n = d["config"]["N"]
    bytes_per_step = d["passes_per_step"] * 4 * n * n
    assert abs(bytes_per_step / (d["ms_per_step"] * 1e-3) / 1e9 - d["value"]) < 5e-3 * d["value"]
    assert d["gpu_launches"] == d["steps"]                       # one persistent kernel per solve
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    if d["n_gpus"] == 1:
        assert d["scaling"] == "weak"
        if d["cpu_baseline"] is not None:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    else:
        assert d["scaling"] == "strong" and d["cpu_baseline"] is None
        assert r["traffic"] is None or d["n_gpus"] == 1
    if d["e2e"] is not None:
        e = d["e2e"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
        assert e["h2d_bytes_per_step"] == 4 * d["config"]["rows_per_gpu"] * n and e["value"] < d["value"]


def test_there_is_a_committed_line_for_every_gpu_count():
    counts = set()
    for path in LINES:
        with open(path) as f:
            counts.add(json.loads(f.read())["n_gpus"])
    assert {1, 2, 4, 8} <= counts


R2 = [p for p in LINES if os.path.basename(p).startswith("r2_")]


@pytest.mark.parametrize("path", R2, ids=[os.path.basename(p) for p in R2])
def test_round_2_lines_carry_a_parity_verdict_and_the_north_star_records(path):
    with open(path) as f:
        d = json.loads(f.read())
    assert d["parity"]["bits_equal"] is True and d["parity"]["checked_against"].startswith("tests/golden/")
    assert d["e2e_pageable"] is None or "pageable" in d["e2e_pageable"]["api"]
    names = [r["workload"] for r in d["north_star"]]
    assert "hilbert-131072" in names and "uniform-131072" in names
    for r in d["north_star"]:
        assert r["n_gpus"] == d["n_gpus"] and r["parity"]["bits_equal"] is True
        assert abs(r["value"] - r["passes_per_step"] * 4.0 * r["N"] ** 2 / (r["ms_per_step"] * 1e-3) / 1e9) < 5e-3 * r["value"]
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(r["clocks"]["reasons"])
        if r["workload"] == "hilbert-131072":
            assert r["rounds"] == 23                     # BASELINE.md section 5's prediction, the oracle's count
