"""Runs the C++ acceptance test (tests/cpp/test_similarity_transform.cpp), which replays the
reference's tests/test.cpp scenario by scenario against the C++ entry points of the B200 build."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_similarity_transform.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_similarity_transform.bin")


def build_cpp_test() -> str:
    from eigen_value_b200 import build
    so = build.build()
    libdir = os.path.dirname(so)
    if (not os.path.exists(EXE) or os.path.getmtime(EXE) < os.path.getmtime(SRC)
            or os.path.getmtime(EXE) < os.path.getmtime(so)):
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", SRC, "-o", EXE, f"-L{libdir}", "-lsimilarity_transform",
               f"-Wl,-rpath,{libdir}"]
        subprocess.run(cmd, check=True)
    return EXE


def test_cpp_test_links_against_the_library():
    # no GPU needed: the reference-named C++ entry points resolve at link time
    assert os.path.exists(build_cpp_test())


@pytest.mark.gpu
def test_reference_cpp_scenarios_pass_on_the_gpu():
    exe = build_cpp_test()
    proc = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout
    assert "all checks passed" in proc.stdout
    assert "[ 4 iterations ]" in proc.stdout and "13 round(s)" in proc.stdout
