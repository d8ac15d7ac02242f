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


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ref_outputs():
    """Outputs of the UNMODIFIED reference classes on seeded inputs (tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through its C-ABI; fails (not skips) if the library cannot be loaded."""
    import ac_dsp_b200 as E
    from ac_dsp_b200 import build as b
    b.build()
    E.load()
    return E


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
