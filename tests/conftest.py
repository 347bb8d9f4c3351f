import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_DIR = os.environ.get("NOBS_TEST_MODEL_DIR", "/tmp/nobs_whisper_models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def model_dir():
    os.makedirs(MODEL_DIR, exist_ok=True)
    return MODEL_DIR


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "hf_micro.npz"))


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative' error the north star's tolerances are stated in)."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
