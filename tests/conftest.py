import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def dense(sparse):
    """Expand the sparse count representation of tests/golden/golden.json."""
    out = np.zeros(sparse["size"], dtype=np.int64)
    out[sparse["idx"]] = sparse["val"]
    return out


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_profiles():
    return np.load(os.path.join(GOLDEN_DIR, "golden_profiles.npz"))


@pytest.fixture(scope="session")
def tutorial_texts():
    texts = {}
    tdir = os.path.join(GOLDEN_DIR, "tutorial")
    for name in sorted(os.listdir(tdir)):
        with open(os.path.join(tdir, name)) as f:
            texts[os.path.splitext(name)[0]] = f.read()
    return texts
