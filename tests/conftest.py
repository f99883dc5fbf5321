import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly rather than silently skip when selected without a device: the product has
    # no CPU path.  When NOT selected (`-m "not gpu"`) pytest deselects them before they run.
    pass


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


@pytest.fixture(scope="session", autouse=True)
def _random_weights():
    """All tests use the seeded random-init weight recipe (no checkpoints offline), with the
    reference's 'resnet' name pointing at ResNet-50 as BASELINE.json's configs do."""
    from i2v_b200 import backbones
    backbones.set_weight_policy("random", seed=0)
    backbones.ARCH_OVERRIDE.update({"resnet": "resnet50", "densenet": "densenet121"})
    yield


def ulp_diff(a, b):
    """Distance in float32 units-in-the-last-place between two float32 arrays (±0 are equal)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)
