import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(d, "golden_v1.npz"))
    with open(os.path.join(d, "golden_v1.json")) as f:
        meta = json.load(f)
    return arrays, meta


@pytest.fixture(scope="session")
def oracle():
    from oracle import hvit_oracle
    return hvit_oracle
