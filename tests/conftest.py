import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU skips the gpu-marked tests instead of erroring in their fixtures.  On a GPU
    box nothing is skipped: a missing libdct_b200.so must fail loudly there (no CPU fallback), not hide as a skip."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    return np.load(path)


@pytest.fixture(scope="session")
def golden_sup():
    """Supervised-branch fixtures (oracle/make_golden_sup.py): CrossEntropyLoss2d + the meter of the same line."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_sup.npz"))


@pytest.fixture(scope="session")
def golden_aux():
    """Class-map / one-hot / ensemble / kappa fixtures (oracle/make_golden_aux.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_aux.npz"))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # oracle/oracle.py (test infrastructure)
    O.build()
    return O
