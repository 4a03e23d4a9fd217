import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# synthetic random-init weights by design (no network): skip the ImageNet download attempt of
# ess_b200.StyleEncoderE2VID; tests/test_host_logic.py covers the attempt + warning path explicitly
os.environ.setdefault('ESS_B200_PRETRAINED', '0')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200, sm_100a)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import torch
    return torch.load(os.path.join(ROOT, 'tests', 'golden', 'ess_tiny.pt'), weights_only=False)
