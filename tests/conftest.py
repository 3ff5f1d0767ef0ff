import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session', autouse=True)
def _built_artifacts():
    """Build libpavgpu.so / _pyrows.so / the oracle library when they are missing or stale (they normally travel with
    the tree; nvcc cross-compiles without a GPU, so this also works on the CPU box)."""
    from oracle import pyoracle
    from pav_b200 import build
    build.build()
    pyoracle.build()
