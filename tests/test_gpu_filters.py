"""GPU parity, row N3: key-frame radius-outlier removal + range box against the CPU restatement (index work: the kept
set and its order must be identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_keyframe_filter_matches_oracle():
    import lsdreg
    from lsdreg import synth
    from oracle.filters import keyframe_filter
    rng = np.random.default_rng(4)
    scan = synth.scan64(7, 700)                                     # ring structure: far points are genuinely isolated
    clutter = np.zeros((3000, 4), np.float32)
    clutter[:, :3] = rng.uniform(-80, 80, (3000, 3)); clutter[:, 2] = rng.uniform(5, 30, 3000)   # floating outliers
    dense = np.zeros((2000, 4), np.float32)
    dense[:, :3] = rng.normal(0, 0.3, (2000, 3)) + [5.0, 5.0, 1.0]  # > 7 points per voxel: overflow levels
    pts = np.concatenate([scan, clutter, dense]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    pts[0, 0] = 0.0                                                 # |x| > 0 is strict: dropped by the range box
    for kw in (dict(radius=1.0, min_neighbors=3, min_range=0.0, max_range=50.0), dict(radius=0.5, min_neighbors=1, min_range=2.0, max_range=1e9),
               dict(radius=0.0, min_neighbors=0, min_range=0.0, max_range=30.0)):
        want = keyframe_filter(pts, **kw)
        got = lsdreg.keyframe_filter(pts, **kw)
        assert 0 < len(want) < len(pts)
        np.testing.assert_array_equal(got, want)
    assert len(lsdreg.keyframe_filter(np.zeros((0, 4), np.float32))) == 0
    lone = np.array([[10.0, 10.0, 1.0, 0.0]], np.float32)
    assert len(lsdreg.keyframe_filter(lone)) == 0                   # a single point has only itself in range
