"""A CPU racecheck of the round kernels (compute-sanitizer is closed on the GPU pool).

The emulation harness (tests/cuda_emu) is rebuilt under ThreadSanitizer with every CUDA thread
announced as a thread of its own and fiber switches NOT synchronising: two CUDA threads touching the
same shared- or global-memory address are then a reported race unless the kernel ordered them through
__syncthreads / __syncwarp / a shuffle / an mbarrier or the atomics of the grid barrier (cuda_emu.h,
TSan section).  A clean run shows that a synchronisation chain exists between all conflicting accesses
the cases exercise -- barrier placement, the parity double-buffering of the row sums, prefetch-slot
reuse, the per-row completion counters, the cluster kernel's remote stores, the cross-GPU exchange.
TSan does not model stand-alone fences, so relaxed atomics are strengthened in this mode: the choice of
fence / scope on the hardware is NOT what is verified here.

The mutants prove the check has teeth: taking away one synchronisation the kernels rely on is reported.
"""
import os
import subprocess
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda_emu"))
import build as emu_build  # noqa: E402

ENV = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66")


def run(exe, *args, timeout=600, halt=False):
    env = dict(ENV, TSAN_OPTIONS=ENV["TSAN_OPTIONS"].replace("halt_on_error=0", "halt_on_error=1")) if halt else ENV
    proc = subprocess.run([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=timeout)
    if "FATAL: ThreadSanitizer" in proc.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + proc.stderr.strip().splitlines()[0])
    return proc


@pytest.fixture(scope="module")
def tsan_exe():
    return emu_build.build_tsan()


def test_positive_control_missing_syncthreads_is_reported(tsan_exe):
    proc = run(tsan_exe, "racy")
    assert "WARNING: ThreadSanitizer: data race" in proc.stderr and "racy_kernel" in proc.stderr


def test_round_kernels_are_race_free(tsan_exe):
    """general loop (both forms, scalar, relative stop), resident-e (prefetch, static / dynamic units, resident
    rows, bf16, two units per row, fp64 accumulation, scalar units for dim % 4 != 0 on one and two GPUs), cluster kernel, wide kernel (one and three windows, 2 GPUs), 2 and 3
    emulated GPUs (tsan_main.cpp)."""
    proc = run(tsan_exe)
    assert "ThreadSanitizer" not in proc.stderr, proc.stderr[:4000]
    assert proc.returncode == 0, proc.stdout
    assert proc.stdout.count("rc=0") == 17 and "agree=0" not in proc.stdout


MUTANTS = [
    # the barrier between rebuilding the eigenvector chunk in shared memory and the rows that read it
    ("general_chunk_barrier", "kernels.cuh",
     "      __syncthreads();\n      // the staged window holds up to four 8192-column chunks",
     "      // the staged window holds up to four 8192-column chunks"),
    # the parity double-buffering of the row-sum vector: one barrier per round is only enough because a CTA
    # that runs one round ahead writes the OTHER buffer
    ("resident_e_single_buffered_s", "kernels_sc.cuh",
     "    float* Scur = p.S[par];",
     "    float* Scur = p.S[0];"),
]


@pytest.mark.parametrize("tag,filename,old,new", MUTANTS, ids=[m[0] for m in MUTANTS])
def test_removing_a_synchronisation_is_reported(tag, filename, old, new):
    exe = emu_build.build_tsan_mutant(tag, filename, old, new)
    proc = run(exe, halt=True)                      # the first report ends the run
    assert "WARNING: ThreadSanitizer: data race" in proc.stderr
