"""The reference-side binding: a consumer compiled against the reference's UNMODIFIED
GpuClauseSharer.h, calling only the abstract class and makeGpuClauseSharerPtr, linked against
libgpushare_b200_shim.so + libgpushare_b200.so (built by __graft_entry__.build() in the
container where /root/reference exists; the binaries travel to the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "gpusharesat_b200", "shim", "shim_selftest")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(EXE), reason="shim_selftest not built")]


def test_cpp_consumer_through_reference_header():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim selftest ok" in r.stdout
