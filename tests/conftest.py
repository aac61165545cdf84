import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _has_gpu():
    """A CUDA device as the library sees it: mchb_create on device 0 (no torch needed); torch's
    answer only where the library cannot be loaded at all."""
    try:
        import ctypes

        from mchap_b200 import _lib

        lib = _lib.load()
        h = ctypes.c_void_p()
        rc = lib.mchb_create(0, ctypes.byref(h))
        if rc == _lib.MCHB_OK:
            lib.mchb_destroy(h)
            return True
        return False
    except Exception:
        pass
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Fixtures written by tests/golden/make_golden.py from the real reference."""

    def __init__(self):
        self._z = np.load(GOLDEN_PATH)
        self.meta = json.loads(bytes(self._z["meta_json"]).decode())

    def __getitem__(self, key):
        return self._z[key]

    def get(self, key, default=None):
        return self._z[key] if key in self._z.files else default

    def __contains__(self, key):
        return key in self._z.files


@pytest.fixture(scope="session")
def golden():
    return Golden()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o
