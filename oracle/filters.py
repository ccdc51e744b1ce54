"""TEST INFRASTRUCTURE ONLY — CPU restatement of the key-frame filters (row N3):
pcl::RadiusOutlierRemoval as configured at slam/src/slam.cpp:104-108 (+ call :398-403) and pointsDistanceFilter
(slam/common/slam_utils.cpp:236-247).  PARITY UNPINNED for the PCL part (PCL 1.9.1 is external): restated from its
published algorithm — a point is kept iff radiusSearch(point, radius), which returns the point itself as well, finds
MORE than min_neighbors points; distances are float32 squared L2, compared strictly with radius^2 (FLANN).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def radius_outlier_mask(pts: np.ndarray, radius: float = 1.0, min_neighbors: int = 3) -> np.ndarray:
    xyz = np.ascontiguousarray(pts[:, :3], np.float32)
    tree = cKDTree(xyz.astype(np.float64))
    r2 = np.float32(radius) * np.float32(radius)
    keep = np.zeros(len(xyz), bool)
    cand = tree.query_ball_point(xyz.astype(np.float64), radius * 1.001)
    for i, c in enumerate(cand):
        d = xyz[c] - xyz[i]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]   # float32, same association as the kernel
        keep[i] = int((d2 < r2).sum()) > min_neighbors
    return keep


def distance_filter_mask(pts: np.ndarray, min_range: float, max_range: float) -> np.ndarray:
    ax, ay = np.abs(pts[:, 0]), np.abs(pts[:, 1])
    return (ax > min_range) & (ax < max_range) & (ay > min_range) & (ay < max_range)


def keyframe_filter(pts: np.ndarray, radius=1.0, min_neighbors=3, min_range=0.0, max_range=1e9) -> np.ndarray:
    """slam.cpp:398-410: outlier removal first, then the range box, order preserved."""
    f = pts[radius_outlier_mask(pts, radius, min_neighbors)] if radius > 0 else pts
    return f[distance_filter_mask(f, min_range, max_range)]
