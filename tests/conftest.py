import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """Without a usable CUDA device the `gpu` tests are skipped instead of failing one by one (the no-fallback contract itself is
    checked on purpose by tests/test_host_cpu.py::test_no_gpu_fails_loudly, which is not a `gpu` test)."""
    if not any('gpu' in it.keywords for it in items):
        return
    try:
        from pav_b200 import _capi
        n = _capi.lib().pavgpu_device_count()
    except Exception:  # noqa: BLE001  (library not built yet: the session fixture builds it; decide then)
        return
    if n <= 0:
        skip = pytest.mark.skip(reason='no usable CUDA device (pavgpu_device_count() <= 0)')
        for it in items:
            if 'gpu' in it.keywords:
                it.add_marker(skip)


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
