import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_goldens.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def state_dict():
    from puzzlenet_b200.weights import synthetic_state_dict
    return synthetic_state_dict(0)


@pytest.fixture(scope="session")
def cuda_model(state_dict):
    """TouchedRegraster on cuda:0 with the synthetic weights (GPU tests only)."""
    import types
    import torch
    from puzzlenet_b200.model5_b import TouchedRegraster
    model = TouchedRegraster(types.SimpleNamespace(dataset="vase"))
    model.load_state_dict(state_dict, strict=True)
    return model.to(torch.device("cuda:0")).eval()
