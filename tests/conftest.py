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


def xform_magnitude(center, scale):
    """Magnitude of the terms of transform_preds (x * s/W + c - s/2, post_transforms.py:6-48) per sample, [N,1,2]:
    |c| + |s * 100|.  Image-space coordinates near 0 are a difference of terms this large, so their f32 rounding
    (one ulp of ~130 is 1.5e-5) is the floor of any element-wise comparison."""
    c = np.abs(np.asarray(center, np.float64)); s = np.abs(np.asarray(scale, np.float64)) * 100.0
    return (c + s)[:, None, :]


def assert_coords_close(a, b, rtol=1e-5, atol=2e-5, what="", mag=None):
    """Float coordinates: 1e-5 relative (north_star) with a 2e-5 px absolute floor for values
    near zero; NaNs must coincide.  mag (optional, broadcastable): magnitude of the terms the coordinate was
    computed from (xform_magnitude for image-space coordinates) — the relative bound is taken against
    max(|b|, mag), element-wise."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{what}: NaN pattern differs"
    ref_mag = np.abs(b) if mag is None else np.maximum(np.abs(b), np.broadcast_to(np.asarray(mag, np.float64), b.shape))
    ok = np.abs(a - b) <= atol + rtol * ref_mag
    ok |= nan_a
    assert ok.all(), f"{what}: max abs diff {np.nanmax(np.abs(a - b))} at {np.argwhere(~ok)[:5]}"


def canon_candidates(c):
    """torch.topk leaves the order of EQUAL values unspecified (and, when the k-th value is tied with values that
    did not make the cut, which of them it returns).  Canonical form for comparing candidate lists [B,N,5]: rows
    sorted by (-confidence, y, x) per image, and a mask of the rows whose confidence is strictly above the k-th
    value — those are the rows every correct top-k must return."""
    c = np.asarray(c).copy()
    det = np.zeros(c.shape[:2], bool)
    for b in range(c.shape[0]):
        order = np.lexsort((c[b, :, 0], c[b, :, 1], -c[b, :, 4].astype(np.float64)))
        c[b] = c[b][order]
        det[b] = c[b, :, 4] > c[b, -1, 4]
    return c, det
