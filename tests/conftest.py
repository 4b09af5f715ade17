import json
import os
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def weight_dir():
    d = os.path.join(tempfile.gettempdir(), "vt_b200_weights")
    os.makedirs(d, exist_ok=True)
    return d


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library and the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as g

    g.build()
    return True
