import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """The C-ABI library is a build artefact (git-ignored).  Build it if a fresh checkout has none:
    nvcc cross-compiles sm_100a without a GPU; the tests never fall back to anything else."""
    import subprocess
    lib = os.path.join(ROOT, "gpusharesat_b200", "libgpushare_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "gpusharesat_b200", "csrc"), "-j8"],
                              stdout=subprocess.DEVNULL)


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
