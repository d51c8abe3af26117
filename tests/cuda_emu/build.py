"""Builds the CPU emulation harness: the round kernels of eigen_value_b200/csrc compiled for the
host against tests/cuda_emu/cuda_emu.h (TEST INFRASTRUCTURE; see that header).

The kernel sources are used as they are, with one mechanical rewrite: CUDA's
`extern __shared__ T name[];` (dynamic shared memory) has no C++ spelling, so each such line becomes
`T* name = reinterpret_cast<T*>(emu::dynamic_smem());` in a scratch copy under _build/.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "eigen_value_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libcuda_emu.so")
KERNEL_SOURCES = ["kernels.cuh", "kernels_sc.cuh", "kernels_wide.cuh", "kernels_cluster.cuh", "launch_plan.hpp"]
OWN_SOURCES = ["cuda_emu.h", "ptx_emu.h", "emu_driver.cpp", os.path.join("fake_include", "cuda_runtime.h"),
               os.path.join("fake_include", "cooperative_groups.h")]

DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];")


def rewrite(text: str) -> str:
    return DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(emu::dynamic_smem());", text)


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in KERNEL_SOURCES] + [os.path.join(HERE, f) for f in OWN_SOURCES] + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False) -> str:
    if not force and not stale():
        return SO
    os.makedirs(BUILD, exist_ok=True)
    for f in KERNEL_SOURCES:
        with open(os.path.join(CSRC, f)) as src:
            text = src.read()
        new = rewrite(text)
        if "extern __shared__" in new:
            raise RuntimeError(f"{f}: an `extern __shared__` declaration was not rewritten")
        with open(os.path.join(BUILD, f), "w") as dst:
            dst.write(new)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-g", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-pthread", "-Wall", "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-Wno-sign-compare",
           '-DST_PTX_HEADER="ptx_emu.h"',
           "-I", os.path.join(HERE, "fake_include"), "-I", BUILD, "-I", HERE,
           "-o", SO, os.path.join(HERE, "emu_driver.cpp")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout[-6000:])
        raise RuntimeError("building the CUDA emulation harness failed")
    return SO


LIB_EMU = os.path.join(HERE, "libsimilarity_transform_emu.so")
LIB_SOURCES = ["solver.cu", "abi.cu", "similarity_transform.hpp"]
RT_SOURCES = ["cuda_runtime_emu.h", "emu_rt.cpp", os.path.join("fake_include_rt", "cuda_runtime.h")]


def build_library(force: bool = False) -> str:
    """The WHOLE library (C ABI + Context::solve + kernels) for the host, on the pretend CUDA runtime of
    cuda_runtime_emu.h -> libsimilarity_transform_emu.so.  Used by tests/conftest.py under ST_EMULATED_LIB=1."""
    build(force=False)
    deps = ([os.path.join(CSRC, f) for f in LIB_SOURCES] + [os.path.join(HERE, f) for f in RT_SOURCES + OWN_SOURCES]
            + [os.path.join(ROOT, "include", "similarity_transform.h"), SO, __file__])
    if not force and os.path.exists(LIB_EMU) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_EMU) for d in deps):
        return LIB_EMU
    rt_dir = os.path.join(BUILD, "library")
    os.makedirs(rt_dir, exist_ok=True)
    for f in KERNEL_SOURCES:
        shutil.copyfile(os.path.join(BUILD, f), os.path.join(rt_dir, f))
    units = []
    for f in LIB_SOURCES:
        with open(os.path.join(CSRC, f)) as src:
            text = src.read()
        text = text.replace('"../../include/similarity_transform.h"',
                            '"' + os.path.join(ROOT, "include", "similarity_transform.h") + '"')
        out = os.path.join(rt_dir, f.replace(".cu", ".cpp"))
        with open(out, "w") as dst:
            dst.write(text)
        if f.endswith(".cu"):
            units.append(out)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [cxx, "-std=c++17", "-O2", "-g", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
           "-pthread", "-w", '-DST_PTX_HEADER="ptx_emu.h"',
           "-I", os.path.join(HERE, "fake_include_rt"), "-I", rt_dir, "-I", HERE,
           "-o", LIB_EMU, *units, os.path.join(HERE, "emu_rt.cpp")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout[-8000:])
        raise RuntimeError("building the emulated library failed")
    return LIB_EMU


TSAN_EXE = os.path.join(HERE, "emu_tsan.bin")


def _compile_tsan(include_dir: str, exe: str) -> str:
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-pthread",
           "-fsanitize=thread", "-fno-omit-frame-pointer", "-w",
           '-DST_PTX_HEADER="ptx_emu.h"',
           "-I", os.path.join(HERE, "fake_include"), "-I", include_dir, "-I", HERE,
           "-o", exe, os.path.join(HERE, "emu_driver.cpp"), os.path.join(HERE, "tsan_main.cpp")]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout[-6000:])
        raise RuntimeError("building the ThreadSanitizer runner failed")
    return exe


def build_tsan(force: bool = False) -> str:
    """The same sources as an executable under ThreadSanitizer (cuda_emu.h, TSan mode): a CPU racecheck."""
    build(force=False)                       # refreshes the rewritten kernel copies when needed
    if not force and os.path.exists(TSAN_EXE) and os.path.getmtime(TSAN_EXE) >= max(
            os.path.getmtime(SO), os.path.getmtime(os.path.join(HERE, "tsan_main.cpp"))):
        return TSAN_EXE
    return _compile_tsan(BUILD, TSAN_EXE)


def build_tsan_mutant(tag: str, filename: str, old: str, new: str) -> str:
    """A TSan runner whose copy of `filename` has `old` replaced by `new` (exactly one occurrence): used
    to show that the racecheck notices when a synchronisation the kernels rely on is taken away."""
    build(force=False)
    mdir = os.path.join(BUILD, "mutant_" + tag)
    os.makedirs(mdir, exist_ok=True)
    for f in KERNEL_SOURCES:
        with open(os.path.join(BUILD, f)) as src:
            text = src.read()
        if f == filename:
            if text.count(old) != 1:
                raise RuntimeError(f"mutation {tag}: expected exactly one occurrence in {filename}, found {text.count(old)}")
            text = text.replace(old, new)
        with open(os.path.join(mdir, f), "w") as dst:
            dst.write(text)
    return _compile_tsan(mdir, os.path.join(mdir, "emu_tsan_mutant.bin"))


if __name__ == "__main__":
    print(build(force=True))
    if "--tsan" in sys.argv:
        print(build_tsan(force=True))
    if "--library" in sys.argv:
        print(build_library(force=True))
