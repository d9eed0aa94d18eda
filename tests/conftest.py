import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def params():
    from ros2_mpc_b200 import load_params
    return load_params()


@pytest.fixture(scope="session")
def built():
    """Native artefacts (libb200mpc.so, oracle) exist; builds them on demand where nvcc/gcc are present."""
    import __graft_entry__ as g
    g.build()
    return True
