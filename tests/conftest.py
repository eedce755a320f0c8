import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_coords_close(a, b, rtol=1e-5, atol=2e-5, what=""):
    """Float coordinates: 1e-5 relative (north_star) with a 2e-5 px absolute floor for values
    near zero; NaNs must coincide."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs"
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    ok |= nan_a
    assert ok.all(), f"{what}: max abs diff {np.nanmax(np.abs(a - b))} at {np.argwhere(~ok)[:5]}"
