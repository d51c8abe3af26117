"""The kernels the committed measurements were taken on are still the kernels that ship.

profiles/r1_measured_kernels_sass_digest.json holds a digest of the instruction stream of every kernel in the build
the round-1 numbers (profiles/) were measured on.  This test disassembles the CURRENT library
and requires the same instruction streams (template parameters added since then only change the mangled names).
If it fails, a measured kernel was changed: re-measure, then refresh the digest (tools/sass_diff.py --write-digest).
Opt-in variants added later (relative stop, bf16 storage, fp64 accumulation) are separate instantiations and not
listed; the standalone find_max / stop kernels were reworked on purpose and are excluded."""
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
    current = sass_diff.digest(sass_diff.split(str(dump)))
    with open(os.path.join(ROOT, "profiles", "r1_measured_kernels_sass_digest.json")) as f:
        doc = json.load(f)
    assert len(doc["kernels"]) == 39
    changed = []
    for name, want in doc["kernels"].items():
        cands = [name] + [name.replace("EEvNS_11RoundParamsE", sfx + "EEvNS_11RoundParamsE")
                          for sfx in doc["mangled_suffixes_added_since"] if sfx]
        got = next((current[c] for c in cands if c in current), None)
        if got is None or got["sha256"] != want["sha256"]:
            changed.append(name)
    assert not changed, "measured kernels changed (re-measure, then refresh the digest): " + ", ".join(changed)
