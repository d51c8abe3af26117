"""The kernels the committed measurements were taken on are still the kernels that ship.

profiles/r2_measured_kernels_sass_digest.json holds a digest of the instruction stream of every round-loop kernel in
the build the round-2 numbers (profiles/r2_*) were measured on; it is refreshed in the tree a GPU
measurement is taken from (`python tools/sass_diff.py --refresh`) and committed together with the numbers.  This test
disassembles the CURRENT library and requires the same instruction streams: if it fails, a measured kernel was
changed after its last measurement -- re-measure, then refresh the digest."""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_diff  # noqa: E402


def test_measured_kernels_are_instruction_identical_in_the_current_build(tmp_path):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from eigen_value_b200 import build
    so = build.build()
    dump = tmp_path / "current.sass"
    with open(dump, "w") as f:
        subprocess.run([cuobjdump, "-sass", so], stdout=f, check=True)
    current = {k: v for k, v in sass_diff.digest(sass_diff.split(str(dump))).items() if "round_loop" in k}
    with open(os.path.join(ROOT, "profiles", "r2_measured_kernels_sass_digest.json")) as f:
        doc = json.load(f)
    assert len(doc["kernels"]) >= 10
    changed = [name for name, want in doc["kernels"].items()
               if name not in current or current[name]["sha256"] != want["sha256"]]
    added = [name for name in current if name not in doc["kernels"]]
    assert not changed and not added, ("round-loop kernels changed since the last committed measurement (re-measure, then "
                                       "`python tools/sass_diff.py --refresh`): " + ", ".join(changed + added))
