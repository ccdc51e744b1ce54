"""Generates tests/golden/lio_ref_small.npz from the COMPILED REFERENCE (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Contents (seeded synthetic inputs; the reference ships no vectors of its own):
  map [M,4], query [Q,3]            inputs
  ivox18_ids/cnt, ivox74_ids/cnt    faster_lio::IVox::GetClosestPoint(k=5, max_sq=5) neighbour ids, sorted
  ikd_ids/d2/cnt                    KD_TREE::Nearest_Search(k=5) ids + the tree's own fp32 distances
  plane_in [P,5,3], plane_abcd/ok   esti_plane<float>(.., 0.1f) on the iVox neighbours (reference row order)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from lsdreg import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

assert O.HAVE_REF, "oracle/_ref/libref_lio.so missing: run `make -C oracle ref` where /root/reference exists"

m = synth.block_map(7, 1, 1, 0.5)[::2].copy()              # ~30 k map points
scan = synth.scan64(8, 60)                                  # ~3.8 k rays
down = O.voxelgrid(scan, 0.5)
R = synth.rot_from_rpy(0.02, -0.01, 0.7)
t = synth.block_center(0, 0) + np.array([2.0, 1.0, 0.0])
q = (down[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)

out = dict(map=m, query=q)
for nearby in (18, 74):
    iv = O.RefIvox(0.5, nearby)
    iv.add(m, 0)
    ids, xyz, cnt = iv.knn(q, 5, 5.0)
    out[f"ivox{nearby}_ids"] = np.sort(ids, 1)
    out[f"ivox{nearby}_cnt"] = cnt
    if nearby == 18:
        sel = cnt >= 5
        out["plane_in"] = xyz[sel]
        pa, ok = O.ref_esti_plane(xyz[sel], 0.1)
        out["plane_abcd"], out["plane_ok"] = pa, ok
kd = O.RefIkd()
kd.build(m)
ids, d2, cnt = kd.knn(q, 5)
out["ikd_ids"], out["ikd_d2"], out["ikd_cnt"] = ids, d2, cnt
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lio_ref_small.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")
