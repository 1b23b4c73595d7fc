import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    data = np.load(path)
    cases = {}
    for row in data["meta"]:
        family, key, kind, n, b, s = str(row).split("|")
        cases.setdefault(family, []).append(dict(key=key, kind=kind, n=int(n), bucket=None if int(b) < 0 else int(b), s=int(s)))
    return data, cases
