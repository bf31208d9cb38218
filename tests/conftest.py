import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gp_ctx():
    """One GPU context for the session; fails loudly (no skip, no fallback) when CUDA or libcngp is missing."""
    from corenav_gp_b200.api import GpContext
    ctx = GpContext(device=0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def slipval():
    import numpy as np
    d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "slipVal.csv"), delimiter=",")
    return d[:, 0], d[:, 1]


@pytest.fixture(scope="session")
def gp_ctx32():
    """FP32-mode context (north_star: 1e-4): variance phase as 3xTF32 on the tensor cores."""
    from corenav_gp_b200.api import GpContext
    ctx = GpContext(device=0, precision="f32")
    yield ctx
    ctx.close()
