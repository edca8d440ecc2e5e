import os
import random
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'needs_reference: needs /root/reference (build container only)')


def _has_gpu():
    try:
        from openfermion_b200 import _lib
        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    have_gpu = None
    have_ref = os.path.isdir('/root/reference/src/openfermion')
    for item in items:
        if 'gpu' in item.keywords:
            if have_gpu is None:
                have_gpu = _has_gpu()
            if not have_gpu:
                item.add_marker(pytest.mark.skip(reason='no CUDA device'))
        if 'needs_reference' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='reference tree not present'))


@pytest.fixture(autouse=True)
def _seed():
    # the reference seeds both generators before every test (conftest.py:32-44)
    random.seed(0)
    numpy.random.seed(0)
