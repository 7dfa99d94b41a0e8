import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


REAL_CONFIGS = ("seird_090_090_10_360", "heat_1_20_05_80_5", "euler_006_200_03_400_6", "seird_120_010_05_480",
                "euler_006_050_01_400_6")


@pytest.fixture
def ctx():
    """Process-wide library context, reset to the reference's RBF kernel family before every test."""
    from gpbo_pkg import pkg

    c = pkg.default_context(0)
    c.set_kernel_family(0)
    return c
