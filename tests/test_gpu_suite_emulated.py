"""The `-m gpu` test files, run on the CPU against the WHOLE library built on the emulation harness.

tests/cuda_emu/build.py::build_library compiles solver.cu + abi.cu + the kernels for the host on a
pretend CUDA runtime (cuda_runtime_emu.h: device memory is host memory, launches are synchronous; the
pretend device gets 148 SMs here, so Context::solve plans exactly the launch shapes it plans on a B200) into tests/cuda_emu/libsimilarity_transform_emu.so; with ST_EMULATED_LIB=1 tests/conftest.py
points the Python binding at it for the test session.  So the C ABI, Context::solve's launch planning,
the Python mirror and the GPU tests' own code are exercised where no GPU exists -- the same assertions
the B200 box runs, on the small and medium cases (the emulated device is orders of magnitude slower).  TEST INFRASTRUCTURE: the
product package never loads that library on its own and still fails loudly without a GPU
(tests/test_abi_symbols.py::test_library_is_sm100a_and_uses_no_cpu_fallback).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_emulated(args, max_dim=4200, timeout=1500):
    env = dict(os.environ, ST_EMULATED_LIB="1", ST_EMU_MAX_DIM=str(max_dim), ST_EMU_DEVICES="4", ST_EMU_SMS="148")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", *args]
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


def test_gpu_test_files_pass_on_the_emulated_library():
    """Every `-m gpu` test file, cases up to N = 4200: the drop-in boundary (golden, 4-byte iter_cnt slot, bad
    arguments, NaN input running to the cap), per-kernel entry points against the reference's unit fixtures, device
    generators, options, both forms, kernel variants, the world == 1 shard path and the in-process sharded solve on
    2 and 4 pretend GPUs, the timeout when a peer never arrives, bit-exact parity with the oracle, relative stop, fp64 accumulation, bf16 and fp8 storage, the streamed solve (block cache, file-backed input), the device group behind one handle (st_group_attach, ST_DEVICES), and 200 random option
    combinations through the launch planner (refused cleanly or bit-exact)."""
    proc = run_emulated(["tests/test_gpu_parity.py", "tests/test_gpu_sharded.py", "tests/test_zz_gpu_bitexact.py",
                         "tests/test_zz_gpu_options_property.py", "tests/test_zzz_gpu_bf16_storage.py",
                         "tests/test_zzz_gpu_fp8_storage.py",
                         "tests/test_zzzz_gpu_streamed.py", "tests/test_zzzz_gpu_group.py"],
                        max_dim=4200)
    assert proc.returncode == 0, proc.stdout[-4000:]
    tail = proc.stdout.strip().splitlines()[-1]
    assert " passed" in tail and "failed" not in tail and "error" not in tail, tail
    assert int(tail.split(" passed")[0].split()[-1]) >= 170, tail
