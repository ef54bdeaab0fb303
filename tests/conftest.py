import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))

    return load


@pytest.fixture(scope="session")
def vgg_weights():
    from nerf_qa_b200 import synth
    return synth.make_vgg_weights(1234)


@pytest.fixture(scope="session")
def alpha_beta():
    import numpy as np
    import torch
    ab = np.load(os.path.join(ROOT, "nerf_qa_b200", "data", "dists_alpha_beta.npy"))
    return torch.from_numpy(ab[0].copy()), torch.from_numpy(ab[1].copy())
