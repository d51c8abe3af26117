"""Drop-in check from the reference's side: its UNMODIFIED Python wrapper
(wrapper/python/similarity_transform.py) and acceptance script (wrapper/python/test.py) are run
from a scratch directory laid out like the reference tree, with `../libsimilarity_transform.so`
pointing at (a) the reference's own C++ built on the CPU SYCL shim and (b) this repo's CUDA
library.  Needs the reference tree, so it runs in the build container only (the GPU box has no
/root/reference; tests/test_gpu_parity.py covers the same calls there through the mirror class)."""
import os
import subprocess
import sys

import pytest

from oracle import ref

REFERENCE = os.environ.get("REFERENCE", "/root/reference")
WRAPPER_DIR = os.path.join(REFERENCE, "wrapper", "python")

pytestmark = pytest.mark.skipif(not os.path.isdir(WRAPPER_DIR), reason="reference tree not present")


def scratch_tree(tmp_path, library):
    """tmp/wrapper/python/{similarity_transform.py,test.py} -> reference files (symlinks),
    tmp/wrapper/libsimilarity_transform.so -> `library`; the wrapper loads '../lib...so' relative
    to its working directory (similarity_transform.py:19)."""
    pydir = tmp_path / "wrapper" / "python"
    pydir.mkdir(parents=True)
    for name in ("similarity_transform.py", "test.py"):
        os.symlink(os.path.join(WRAPPER_DIR, name), pydir / name)
    os.symlink(library, tmp_path / "wrapper" / "libsimilarity_transform.so")
    return pydir


def run_in(pydir, code, timeout=600):
    return subprocess.run([sys.executable, "-c", code], cwd=pydir, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=timeout)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_reference_wrapper_and_acceptance_script_on_the_shim_build(tmp_path):
    pydir = scratch_tree(tmp_path, ref.SO)
    # a smaller DIM than test.py's 1024 keeps the fiber-emulated run short; same code path
    code = ("import test, similarity_transform as st\n"
            "test.DIM = 1 << 8\n"
            "test.main()\n")
    proc = run_in(pydir, code)
    assert proc.returncode == 0, proc.stdout
    assert proc.stdout.count("passed randomized test against 256 x 256 similarity transform") == 4


def test_reference_wrapper_binds_the_cuda_library(tmp_path):
    from eigen_value_b200 import _lib, build
    so = build.build()
    pydir = scratch_tree(tmp_path, so)
    have_gpu = _lib.load().st_device_count() > 0
    if have_gpu:
        code = ("import test\n"
                "test.main()\n")
        proc = run_in(pydir, code)
        assert proc.returncode == 0, proc.stdout
        assert proc.stdout.count("passed randomized test against 1024 x 1024 similarity transform") == 4
    else:
        # no GPU: every symbol the wrapper needs resolves, make_queue leaves the handle NULL and the
        # wrapper raises exactly as it would for a missing SYCL device (similarity_transform.py:39-40)
        code = ("import similarity_transform as st\n"
                "try:\n"
                "    st.EigenValue()\n"
                "except Exception as e:\n"
                "    print('RAISED', e)\n")
        proc = run_in(pydir, code)
        assert proc.returncode == 0, proc.stdout
        assert "RAISED failed to get default SYCL queue" in proc.stdout
