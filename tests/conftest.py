import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def synth():
    from lsd_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def lsd():
    """The product binding.  Fails loudly if the CUDA library is missing (no fallback)."""
    import lsd_b200
    lsd_b200.load()
    return lsd_b200
