import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def mag():
    """The reference's 305x305 synthetic anomaly (tests/test_mag_data.txt) + extent."""
    g = load_golden("mag_grid.npz")
    return g["d"], tuple(float(v) for v in g["extent"])


@pytest.fixture(scope="session")
def oracle():
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    if not os.path.exists(os.path.join(odir, "libradon_oracle.so")):
        subprocess.check_call(["make", "-C", odir, "libradon_oracle.so"])
    import curie_oracle
    return curie_oracle


TAPERS = {"hanning": np.hanning, "hamming": np.hamming, "none": None,
          "blackman": np.blackman, "kaiser": np.kaiser}
