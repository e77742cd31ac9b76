import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass without a device: outside `-m gpu` runs they are skipped,
    inside them a missing device is an error raised by the `cuda` fixture."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return  # explicit GPU run on a box without a GPU: let the tests fail loudly
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cuda():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    from numbotics_b200 import _cabi
    _cabi.load()   # raises if libnbx.so is missing: no silent fallback
    return torch.device("cuda", 0)


@pytest.fixture(scope="session")
def golden_kinova():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "fkjac_kinova.npz")))


@pytest.fixture(scope="session")
def golden_raw():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "helpers_raw.npz")))


@pytest.fixture(scope="session")
def golden_serial30():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "fkjac_serial30.npz")))
