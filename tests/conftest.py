import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def fixture_images():
    d = np.load(os.path.join(ROOT, "tests", "golden", "stereo_fixture.npz"))
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def cv2_vectors():
    d = np.load(os.path.join(ROOT, "tests", "golden", "cv2_vectors.npz"))
    return {k: d[k] for k in d.files}
