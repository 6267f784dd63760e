import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: without one (the build container) they are skipped instead of erroring.
    On a box WITH a GPU nothing is skipped -- a missing / unloadable librefnerf_b200.so must fail loudly there."""
    try:
        import torch
        ok = torch.cuda.is_available()
    except Exception:   # noqa: BLE001
        ok = False
    if ok:
        return
    skip = pytest.mark.skip(reason='gpu test: no CUDA device in this environment')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)
