import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the native pieces are built in-tree by __graft_entry__.build(); build them here if a fresh checkout has none
    # (nvcc cross-compiles without a GPU).  This only builds -- the product still fails loudly when it cannot run.
    import subprocess
    lib = os.path.join(ROOT, "slam.jl_b200", "csrc", "libslamklt.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.dirname(lib), "-j4", "libslamklt.so"], stdout=subprocess.DEVNULL)
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.dirname(orc), "liboracle.so"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def ctx():
    import slamklt
    c = slamklt.Context(0)  # raises loudly when the library or the device is missing
    yield c
    c.close()
