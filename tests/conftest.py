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
def golden_mproduct():
    return np.load(os.path.join(GOLDEN, "mproduct.npz"))


@pytest.fixture(scope="session")
def golden_chess():
    return np.load(os.path.join(GOLDEN, "chess6.npz"))


@pytest.fixture(scope="session")
def golden_models():
    return np.load(os.path.join(GOLDEN, "models.npz"))


@pytest.fixture(scope="session")
def golden_minv():
    return np.load(os.path.join(GOLDEN, "minv.npz"))


@pytest.fixture(scope="session")
def golden_data():
    return np.load(os.path.join(GOLDEN, "data_helpers.npz"))


MPRODUCT_CASES = ["t8n50b3", "t8n50b3norm", "t12n33b20", "t5n17b1", "t9n40raw"]
