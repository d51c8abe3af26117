"""Drop-in check from the reference's side: its UNMODIFIED Python wrapper
(wrapper/python/similarity_transform.py) and acceptance script (wrapper/python/test.py) are run
from a scratch directory laid out like the reference tree, with `../libsimilarity_transform.so`
pointing at (a) the reference's own C++ built on the CPU SYCL shim and (b) this repo's CUDA
library.  The two Python files come from the reference tree where it exists (the build container) and
otherwise from oracle/_ref/wrapper/python/ -- byte-for-byte copies made by `make -C oracle ref`
(git-ignored build outputs like the rest of oracle/_ref, which travel to the GPU box with the snapshot).
The `gpu` tests at the bottom are the real drop-in call: reference wrapper/python/test.py:8-18, unmodified,
on a B200 against libsimilarity_transform.so, with ST_DEVICES unset and with ST_DEVICES=all."""
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
WRAPPER_DIR = os.path.join(REFERENCE, "wrapper", "python")
if not os.path.isdir(WRAPPER_DIR):
    WRAPPER_DIR = os.path.join(ROOT, "oracle", "_ref", "wrapper", "python")

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(WRAPPER_DIR, "test.py")),
                                reason="neither the reference tree nor oracle/_ref/wrapper is present")


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


def run_in(pydir, code, timeout=600, env=None):
    return subprocess.run([sys.executable, "-c", code], cwd=pydir, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True, timeout=timeout, env=env)


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


def test_travelling_copies_are_the_reference_files():
    """oracle/_ref/wrapper/python/* must be the reference's files byte for byte (checked where both exist)."""
    ref_dir = os.path.join(REFERENCE, "wrapper", "python")
    copy_dir = os.path.join(ROOT, "oracle", "_ref", "wrapper", "python")
    if not (os.path.isdir(ref_dir) and os.path.isdir(copy_dir)):
        pytest.skip("needs the reference tree and the oracle/_ref copies")
    for name in ("similarity_transform.py", "test.py"):
        with open(os.path.join(ref_dir, name), "rb") as a, open(os.path.join(copy_dir, name), "rb") as b:
            assert a.read() == b.read(), name


_ACCEPT = ("import test\n"
           "test.main()\n")
_PASSED = "passed randomized test against 1024 x 1024 similarity transform"


@pytest.mark.gpu
def test_unmodified_reference_acceptance_script_on_the_cuda_library(tmp_path):
    """reference wrapper/python/test.py:8-18 through wrapper/python/similarity_transform.py:18-78, both
    unmodified, bound to the CUDA library by the relative path the wrapper itself uses (:19)."""
    from eigen_value_b200 import build
    pydir = scratch_tree(tmp_path, build.build())
    env = {k: v for k, v in os.environ.items() if k != "ST_DEVICES"}
    proc = run_in(pydir, _ACCEPT, env=env)
    assert proc.returncode == 0, proc.stdout
    assert proc.stdout.count(_PASSED) == 4, proc.stdout


@pytest.mark.gpu
def test_unmodified_reference_wrapper_with_every_gpu_behind_the_handle(tmp_path):
    """Same script with ST_DEVICES=all: make_queue binds every GPU of the box to the one handle the wrapper
    knows; 1024 rows is below the default sharding threshold, so the second run lowers it (ST_GROUP_MIN_DIM)
    and the acceptance criterion A.v ~ lambda.v is met by the SHARDED solve.  On a one-GPU box the group is
    the handle's own GPU and the call must still pass."""
    from eigen_value_b200 import build
    pydir = scratch_tree(tmp_path, build.build())
    for extra in ({}, {"ST_GROUP_MIN_DIM": "256"}):
        env = dict(os.environ, ST_DEVICES="all", **extra)
        proc = run_in(pydir, _ACCEPT, env=env)
        assert proc.returncode == 0, proc.stdout
        assert proc.stdout.count(_PASSED) == 4, proc.stdout
