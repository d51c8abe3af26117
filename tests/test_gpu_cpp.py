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


def test_cpp_acceptance_on_the_emulated_library():
    """The same C++ program linked against the library built on the CPU emulation harness (tests/cuda_emu, TEST
    INFRASTRUCTURE), two pretend GPUs: the reference's tests/test.cpp scenarios, the streamed solve and the device
    group behind max_eigen_value run through the C++ entry points where no GPU exists."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
    import build as emu_build
    lib = emu_build.build_library()
    exe = os.path.join(ROOT, "tests", "cuda_emu", "cpp_acceptance_emu.bin")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", SRC, "-o", exe, f"-L{os.path.dirname(lib)}", f"-l:{os.path.basename(lib)}",
           f"-Wl,-rpath,{os.path.dirname(lib)}"]
    subprocess.run(cmd, check=True)
    env = dict(os.environ, ST_EMU_DEVICES="2", ST_EMU_SMS="8")
    proc = subprocess.run([exe], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout
    assert "all checks passed" in proc.stdout and "[ 4 iterations ]" in proc.stdout and "13 round(s)" in proc.stdout
    assert "streamed solve: same bits" in proc.stdout and "device group of 2 GPUs" in proc.stdout
