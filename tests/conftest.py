import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "refpin: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def synth_sd():
    """synthetic_weights(seed=0) rebuilt bit-identically from the committed BN statistics."""
    from oracle import synth_weights as sw
    return sw.synthetic_weights(0, 256, bn_stats=sw.load_bn_stats(os.path.join(GOLD, "bn_stats_seed0.npz")))
