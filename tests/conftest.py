import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (no-op when they are up to date)."""
    import __graft_entry__ as g
    if os.path.isdir("/root/reference/src") or not os.path.exists(os.path.join(ROOT, "amici_b200", "lib", "libamici_b200.so")):
        g.build_cuda()
    g.build_cpu_emul()
    return True
