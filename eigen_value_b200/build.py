"""Build recipe: eigen_value_b200/csrc/*.cu -> eigen_value_b200/libsimilarity_transform.so.

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored and travels to the GPU
box with the tree.  `python -m eigen_value_b200.build` rebuilds unconditionally.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_NAME = "libsimilarity_transform.so"   # the artefact name of reference Makefile:69
SO_PATH = os.path.join(HERE, SO_NAME)
SOURCES = ["solver.cu", "abi.cu"]
HEADERS = ["kernels.cuh", "ptx.cuh", "launch_plan.hpp", "kernels_sc.cuh", "kernels_wide.cuh", "kernels_cluster.cuh", "similarity_transform.hpp", os.path.join("..", "..", "include", "similarity_transform.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "--threads", "2",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO_PATH
    cmd = [nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", SO_PATH, *[os.path.join(CSRC, f) for f in SOURCES]]
    env = dict(os.environ)
    # this image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc wants plain g++
    cmd += ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building " + SO_NAME)
    return SO_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
