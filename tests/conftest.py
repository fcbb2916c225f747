import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the engine + oracle once per session if the artefacts are missing or stale."""
    import __graft_entry__ as g
    g.build()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack(bits, n):
    return np.unpackbits(bits, axis=1)[:, :n]
