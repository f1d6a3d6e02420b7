import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "2d-lb_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built library (it is git-ignored): build it in-tree, as __graft_entry__.build() does.
    Only when it is MISSING -- an existing one is what the tests are about, stale or not."""
    from lb_b200 import build
    if not os.path.exists(build.LIB) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        build.build_library(force=True)


def _gpu_expected():
    """True where a GPU is supposed to be present: GPU tests must then FAIL, not skip, if the
    CUDA library cannot see it."""
    if os.path.exists("/dev/nvidia0"):
        return True
    smi = shutil.which("nvidia-smi")
    if smi:
        try:
            return subprocess.run([smi, "-L"], capture_output=True, timeout=20).returncode == 0
        except Exception:
            return False
    return False


@pytest.fixture(scope="session")
def gpu():
    """Builds nothing, substitutes nothing: loads the in-tree CUDA library or fails."""
    from lb_b200 import native
    n = native.lib().lb_device_count()
    if n == 0:
        if _gpu_expected():
            pytest.fail("a GPU is present but liblb_d2q9.so sees no CUDA device")
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    return n


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle
