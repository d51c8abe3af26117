"""bench.py contract checks that need no GPU: the reference arm (CPU oracle port) runs and prints
one JSON line with the agreed keys; non-zero ranks of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    proc = subprocess.run([sys.executable, BENCH, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=600, env=e)
    return proc


def test_reference_arm_prints_one_json_line():
    proc = run_bench("--impl", "reference", "--workload", "hilbert-512", "--steps", "2", "--warmup", "1")
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "hilbert-512" and d["config"]["passes_per_step"] == 13   # 12 rounds + 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "hilbert-512" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    proc = run_bench("--impl", "reference", "--gpus", "2", "--workload", "hilbert-512", "--steps", "1",
                     env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_unknown_workload_is_rejected():
    proc = run_bench("--impl", "reference", "--workload", "lehmer-64")
    assert proc.returncode != 0
