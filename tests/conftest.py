import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference, produced by tests/golden/make_golden.py."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")))


@pytest.fixture(scope="session")
def golden_large():
    """eps at the 256x256-spectrogram geometries (S = 128 from the unmodified reference, S = 256 from the query-chunked
    oracle), produced by tests/golden/make_golden_large.py."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_large.npz")))
