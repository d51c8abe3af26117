"""bench.py's B200 arm, executed end to end on the emulated library (tests/cuda_emu/run_bench_emulated.py):
the contract keys of the JSON line, the e2e leg through max_eigen_value, the Hilbert sweep table, the CPU
baseline and the opt-in switches are exercised where no GPU exists.  Values are meaningless here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "cuda_emu", "run_bench_emulated.py")


def run(*args, timeout=900):
    proc = subprocess.run([sys.executable, RUNNER, *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=timeout)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, proc.stdout[-2000:]                       # ONE JSON line
    return json.loads(lines[0])


def test_default_shape_of_the_line_with_e2e_sweep_table_and_cpu_baseline():
    # a workload smaller than the sweep's largest size: the sweep must not reuse the workload's output vector
    line = run("--workload", "hilbert-256", "--steps", "2", "--warmup", "1", "--scale-base-dim", "512")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["gpu_launches"] == 2
    assert line["rounds"] == 10 and line["config"]["workload"] == "hilbert-256"           # reference README.md:71
    assert line["config"]["stop"] == "absolute" and line["config"]["storage"] == "f32" and line["config"]["accumulate"] == "f32"
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert line["e2e"]["h2d_bytes_per_step"] == 4 * 256 * 256 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["e2e_pageable"]["h2d_bytes_per_step"] == 4 * 256 * 256 and "pageable" in line["e2e_pageable"]["api"]
    assert line["parity"] == {"checked_against": None, "bits_equal": None} and "north_star" not in line   # no golden entry for 256
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert [r["rounds"] for r in line["hilbert_sweep"]] == [9, 10, 12, 13, 14, 15, 17]      # reference README.md:70-76
    assert line["strong_scaling_base"]["workload"] == "hilbert-512" and line["strong_scaling_base"]["rounds"] == 12
    assert line["cpu_baseline"]["sequential_model"]["rounds_main_py"] == 11                # main.py on Hilbert 1024


@pytest.mark.parametrize("extra,rounds", [(("--stop", "relative", "--eps", "1e-6"), None), (("--storage", "bf16"), 13),
                                          (("--accumulate", "f64"), 13)])
def test_opt_in_switches(extra, rounds):
    line = run("--workload", "hilbert-1024", "--steps", "2", "--no-cpu-baseline", "--no-sweep-table", *extra)
    assert "strong_scaling_base" not in line
    if rounds is not None:
        assert line["rounds"] == rounds
    if "--storage" in extra:
        assert line["e2e"] is None and line["roofline"]["bytes_per_launch"] == line["passes_per_step"] * 2 * 1024 * 1024
        assert "bf16 storage" in line["metric"]


def test_parity_verdict_against_the_cpu_computed_expected_bits():
    # hilbert-16384 is in tests/golden/generated_expected.json but far too slow here; the verdict logic itself:
    sys.path.insert(0, ROOT)
    import bench
    e, src = bench.expected_result("hilbert-131072", 1000)
    assert src == "tests/golden/generated_expected.json" and e["iter_count"] == 23
    assert bench.expected_result("uniform-131072", 1000) == (None, None)          # computed for a 50-round cap only
    assert bench.expected_result("uniform-131072", 50)[0]["iter_count"] == 50
    assert bench.expected_result("uniform-65536", 1000)[1] == "tests/golden/gpu_recorded.json"
    assert bench.expected_result("hilbert-32768", 10) == (None, None)             # a cap below the 20 rounds it takes
    ok = bench.parity_record("hilbert-131072", 1000, e["eigen_val"], 23)
    assert ok["bits_equal"] is True and ok["checked_against"].endswith("generated_expected.json")
    assert bench.parity_record("hilbert-131072", 1000, e["eigen_val"], 22)["bits_equal"] is False
    assert bench.parity_record("hilbert-131072", 1000, 2.7381439, 23)["bits_equal"] is False
    assert bench.parity_record("hilbert-777", 1000, 1.0, 3) == {"checked_against": None, "bits_equal": None}


def test_streamed_bench_tool_on_the_emulated_library():
    # tools/bench_streamed.py (streamed solve vs in-device solve of the same matrix), same trick as the runner above
    code = ("import os, sys; sys.path.insert(0, 'tests/cuda_emu'); os.environ.setdefault('ST_EMU_SMS', '8');"
            "import build as b; from eigen_value_b200 import _lib;"
            "_lib._build.SO_PATH = b.build_library(); _lib._build.stale = lambda: False;"
            "sys.path.insert(0, 'tools'); import bench_streamed;"
            "sys.argv = ['bench_streamed.py', '--dim', '512', '--cached', '0.4', '--block-rows', '32', '--steps', '1'];"
            "sys.exit(bench_streamed.main())")
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["bit_identical_to_in_device_solve"] is True and line["rounds"] == 12          # reference README.md:72
    plan = line["plan"]
    assert plan["streamed"] == 1 and plan["blocks"] == 16 and plan["slots"] == 6
    assert plan["h2d_bytes_per_round"] == 10 * 32 * 512 * 4 and plan["h2d_bytes_first"] == 4 * 512 * 512


def test_group_bench_tool_on_the_emulated_library():
    # tools/bench_group.py (drop-in call on one GPU vs a device group), 4 pretend GPUs
    code = ("import os, sys; sys.path.insert(0, 'tests/cuda_emu'); os.environ.setdefault('ST_EMU_SMS', '8');"
            "os.environ['ST_EMU_DEVICES'] = '4';"
            "import build as b; from eigen_value_b200 import _lib;"
            "_lib._build.SO_PATH = b.build_library(); _lib._build.stale = lambda: False;"
            "sys.path.insert(0, 'tools'); import bench_group;"
            "sys.argv = ['bench_group.py', '--dim', '512', '--steps', '1', '--min-dim', '256'];"
            "sys.exit(bench_group.main())")
    proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                          timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["bit_identical"] is True and line["rounds"] == 12 and line["group"]["gpus"] == 4


def test_upload_bench_tool_on_the_emulated_library():
    # tools/bench_upload.py's per-setting child (pageable matrix, ST_UPLOAD_THREADS read at handle creation)
    code = ("import os, sys; sys.path.insert(0, 'tests/cuda_emu'); os.environ.setdefault('ST_EMU_SMS', '8');"
            "import build as b; from eigen_value_b200 import _lib;"
            "_lib._build.SO_PATH = b.build_library(); _lib._build.stale = lambda: False;"
            "sys.path.insert(0, 'tools'); import bench_upload;"
            "sys.exit(bench_upload.child(3000, 1))")
    for threads, staged in (("0", 0), ("3", 2 * 4 * 3000 * 3000)):      # two pageable solves; the pinned ones go direct
        proc = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, ST_UPLOAD_THREADS=threads),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert proc.returncode == 0, proc.stderr[-3000:]
        line = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
        assert line["upload_threads"] == int(threads) and line["staged_bytes"] == staged and line["rounds"] == 15
