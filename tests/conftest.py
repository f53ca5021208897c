import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / 'pets-face-recognition_b200'
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _gpu_ready():
    """`-m gpu` tests are skipped (not errors) on a host without a CUDA device.  On a GPU box nothing is skipped: a missing
    libb200fe.so must fail loudly there (there is no fallback path to hide behind)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False, 'no CUDA device'
    except Exception as e:  # pragma: no cover
        return False, f'torch unavailable: {e}'
    return True, ''


def pytest_collection_modifyitems(config, items):
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=f'gpu test: {why}')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return ROOT / 'tests' / 'golden'
