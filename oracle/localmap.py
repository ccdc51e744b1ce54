"""TEST INFRASTRUCTURE ONLY — numpy restatement of the local-map assembly of
Localization::runUpdateLocalMap (slam/localization/src/localization.cpp:303-373).  The radius search over key-frame
positions is pcl::KdTreeFLANN::radiusSearch (float squared distances, sorted ascending): PARITY UNPINNED for that
external part; the selection rule, the caps and the VoxelGrid call are restated from the reference text."""
from __future__ import annotations

import numpy as np

from . import oracle as O


class OracleLocalMap:
    def __init__(self, resolution=0.5, key_frame_distance=1.0):
        self.res = max(resolution, 0.1)
        self.kfd = np.float32(key_frame_distance)
        self.frames, self.pos = [], []

    def add_keyframe(self, pts, position):
        self.frames.append(np.ascontiguousarray(pts, np.float32)); self.pos.append(np.asarray(position, np.float64))

    def update(self, pose_xyz):
        """-> (cloud or None, n_keyframes_in_radius)"""
        if not self.pos:
            return None, 0
        d = np.asarray(self.pos).astype(np.float32) - np.asarray(pose_xyz, np.float64).astype(np.float32)
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        idx = [i for i in np.argsort(d2, kind="stable") if d2[i] < np.float32(30.0 * 30.0)]
        if not idx:
            return None, 0
        parts, total, accum = [], 0, np.float32(0)
        for i in idx:
            dist = np.sqrt(d2[i])
            if total > 0 and (dist - accum) < self.kfd:
                continue
            accum = dist
            parts.append(self.frames[i]); total += len(self.frames[i])
            if total >= 200000:
                break
        cloud = O.voxelgrid(np.concatenate(parts), self.res)
        if d2[idx[0]] >= 400:
            return None, len(idx)
        return cloud, len(idx)
