import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def known():
    with open(os.path.join(GOLDEN, "reference_known_answers.json")) as fh:
        return json.load(fh)


def read_asc(name):
    """'% N' header then one value per line (reference notebooks/data/*.asc)."""
    with open(os.path.join(GOLDEN, name)) as fh:
        return np.array([float(t) for t in fh.read().split("\n") if t.strip() and not t.startswith("%")])


@pytest.fixture(scope="session")
def nagata_H():
    from oracle.basis import halfbox_symmetries
    sx, sy, sz, tx, tz = halfbox_symmetries()
    return [sx * sy * sz, sz * tx * tz]
