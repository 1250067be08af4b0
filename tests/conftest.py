import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected on a box without a GPU; they are only
    # deselected through -m "not gpu".
    pass


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure the native library and the C oracle exist (cheap no-op when up to date)."""
    import __graft_entry__ as entry

    from octreelib_b200 import _native

    if not os.path.exists(_native.LIB_PATH):
        entry.build()
    from oracle import ransac as oransac

    oransac.build()
    yield


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name + ".npz"))
