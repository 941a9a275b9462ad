import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


@pytest.fixture(scope="session")
def cfg():
    from oracle.arch import ModelConfig
    return ModelConfig()


@pytest.fixture(scope="session")
def sd(cfg):
    """The synthetic reference-layout checkpoint (seed 0) the golden vectors were made with."""
    from oracle.synth import synth_state_dict
    return synth_state_dict(cfg, seed=0)


@pytest.fixture(scope="session")
def golden_src():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "swin_src_golden.npz"))


@pytest.fixture(scope="session")
def golden_app():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "swin_app_golden.npz"))
