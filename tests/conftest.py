import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_modules():
    import numpy as np

    return dict(np.load(os.path.join(GOLDEN, "modules.npz")))


@pytest.fixture(scope="session")
def golden_models():
    import numpy as np

    return dict(np.load(os.path.join(GOLDEN, "models.npz")))


@pytest.fixture(scope="session")
def golden_topology():
    import json

    return json.load(open(os.path.join(GOLDEN, "topology.json")))


@pytest.fixture(scope="session")
def smpl():
    import numpy as np

    return dict(np.load(os.path.join(ROOT, "hm_vae_b200", "data", "smpl24.npz")))
