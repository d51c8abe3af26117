"""Runs bench.py's B200 arm on the emulated library (TEST INFRASTRUCTURE, tests/test_bench_emulated.py).

bench.py needs torch.cuda only as plumbing (an L2-flush buffer, pinned host memory, a scalar all-reduce
buffer); here those calls are pointed at CPU tensors and the Python binding at
tests/cuda_emu/libsimilarity_transform_emu.so, so that every line of bench.py's main() executes -- the
timed loop, the e2e leg through max_eigen_value, the Hilbert sweep, the CPU baseline, the JSON line.
The numbers it prints are meaningless; the point is that the script and its contract keys work.

    python tests/cuda_emu/run_bench_emulated.py [bench.py arguments...]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
os.environ.setdefault("ST_EMU_SMS", "8")

import build as emu_build  # noqa: E402
from eigen_value_b200 import _lib  # noqa: E402

_lib._build.SO_PATH = emu_build.build_library()
_lib._build.stale = lambda: False

import torch  # noqa: E402

torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
_empty, _tensor = torch.empty, torch.tensor
torch.empty = lambda *a, **k: _empty(*a, **{kk: v for kk, v in k.items() if kk != "device"})
torch.tensor = lambda *a, **k: _tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})
torch.Tensor.pin_memory = lambda self, *a, **k: self

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv = ["bench.py"] + sys.argv[1:]
    sys.exit(bench.main())
