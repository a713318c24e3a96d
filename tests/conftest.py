import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        if os.environ.get("NND_TCS_MAP"):          # A/B runs of the whole suite with the streaming kernel's other halo copy mapping
            from ctypes import c_int
            from nndetection_b200 import _lib as L
            L.lib().nnd_conv_set_tcs_map(c_int(int(os.environ["NND_TCS_MAP"])))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
