import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


EMULATED = os.environ.get("ST_EMULATED_LIB") == "1"
EMU_MAX_DIM = int(os.environ.get("ST_EMU_MAX_DIM", "4200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run under gpurun)")
    if EMULATED:
        # TEST-ONLY: exercise the `-m gpu` test files on the CPU against the whole library built on the
        # emulation harness (tests/cuda_emu: C ABI + Context::solve + kernels on a pretend 4-SM device).
        # The product package never does this on its own; it is a way to run the GPU tests' own code, the
        # ABI and the launch planning where no GPU exists.  Results are bit-identical to the hardware's
        # by construction of the kernels' fixed evaluation order, timings are meaningless.
        sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
        import build as emu_build
        from eigen_value_b200 import _lib
        _lib._build.SO_PATH = emu_build.build_library()
        _lib._build.stale = lambda: False


def pytest_report_header(config):
    if EMULATED:
        return ("*** ST_EMULATED_LIB=1: the `-m gpu` tests of this session run on the CPU EMULATION of the library "
                "(tests/cuda_emu), NOT on a GPU -- results say nothing about hardware ***")
    return None


def _have_gpu() -> bool:
    try:
        from eigen_value_b200 import _lib
        return _lib.load().st_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def solver():
    from eigen_value_b200 import Solver
    s = Solver(0)
    yield s
    s.close()


def pytest_collection_modifyitems(config, items):
    # GPU tests never silently pass on a box without a GPU: they are skipped with a reason
    # here (CPU container) and run for real under `-m gpu` on the B200 box.
    if EMULATED:
        # the emulated device is ~1000x slower than a B200: keep the small cases
        slow = pytest.mark.skip(reason=f"emulated library: case larger than ST_EMU_MAX_DIM={EMU_MAX_DIM} or multi-process")
        for item in items:
            params = getattr(getattr(item, "callspec", None), "params", {})
            dims = [v for k, v in params.items() if k in ("dim", "world") and isinstance(v, int)]
            case = params.get("case")
            if isinstance(case, dict) and isinstance(case.get("dim"), int):
                dims.append(case["dim"])
            big_by_name = any(t in item.name for t in ("full_size", "16384", "32768", "beyond_the_resident_limit",
                                                       "many_rounds_beyond", "fused_exchange_multi_gpu", "cpp_acceptance",
                                                       "converges_where_the_reference_test_cannot", "every_kernel_family",
                                                       "general_loop_on_bf16_storage",
                                                       # need a real CUDA tensor / link the real library by name
                                                       "zero_copy_torch", "reference_cpp_scenarios"))
            if big_by_name or any(d > EMU_MAX_DIM for d in dims if d > 8):
                item.add_marker(slow)
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
