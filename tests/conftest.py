import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liboracle.so), built on demand with gcc."""
    from oracle import oracle

    oracle.build()
    oracle.lib()
    return oracle


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cpp_demo():
    """examples/cpp_host/splat_demo, the driver of the C++ host side (include/splat_pipeline.hpp).  Rebuilt with
    g++ when stale (seconds); a binary that travelled with the snapshot is used as it is if that fails."""
    import subprocess

    d = os.path.join(ROOT, "examples", "cpp_host")
    exe = os.path.join(d, "splat_demo")
    if not os.path.exists(os.path.join(ROOT, "splat_b200", "libsplat_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "splat_b200", "csrc")])
    r = subprocess.run(["make", "-s", "-C", d], capture_output=True, text=True)
    if r.returncode != 0 and not os.path.exists(exe):
        pytest.fail("cannot build examples/cpp_host/splat_demo:\n" + r.stderr)
    return exe
