"""GPU: the flat k-NN shape (csrc/knn_flat.cuh — a warp owns 32 queries and walks the ballot-compacted list of voxels
that exist) against the two validated shapes and the oracle, batched (lsd_knn_set_shape) and inside the LIO scan
(lsd_lio_set_knn_shape).  Bar: bit-identical ids, fp32 d2 and counts; bit-identical Nearest_Points and pose.

STATUS: written after this round's GPU budget was spent — it has never run on a GPU.  The flat shape is OFF by default
(shape 0 = the validated kernels), so nothing else depends on it.  Runs in a subprocess, sorts last, NON-STRICT xfail:
it reports xpassed / xfailed and cannot turn the validated suite red.  Round 2 runs it first (tools/knn_shapes_probe.py
times the three shapes on the bench map) and removes the marker.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import lsdreg
from lsdreg import synth
from oracle import oracle as O
from oracle import eskf

lsdreg.init(0)
rng = np.random.default_rng(21)
m = synth.block_map(1, 2, 2, 0.5)
g = lsdreg.HashVoxelMap(0.5, 20)
g.insert(m, 0)
o = O.OracleIvox(0.5, 18, 1 << 18)
o.add(m, 0)

def shapes(gm, q, k, nearby):
    out = {}
    for name, s in (("warp", 1), ("thread", 2), ("flat", 3)):
        gm.set_knn_shape(s)
        out[name] = gm.knn(q, k=k, max_sq=5.0, stencil=nearby)
    gm.set_knn_shape(0)
    return out

# 1. batches of awkward sizes (not a multiple of 32 / 128), every fixed stencil, k = 1 and 5, some queries far from any voxel
for nearby, k, nq in ((18, 5, 70001), (74, 5, 20011), (18, 1, 33333), (6, 5, 4097), (26, 5, 31), (0, 5, 1000)):
    q = m[rng.integers(0, m.shape[0], nq)].copy()
    q[:, :3] += rng.normal(0, 0.15, (nq, 3)).astype(np.float32)
    q[: min(100, nq // 3), :3] += 500.0
    r = shapes(g, q, k, nearby)
    for j in range(3):
        np.testing.assert_array_equal(r["flat"][j].view(np.int32), r["warp"][j].view(np.int32), err_msg=f"flat vs warp {nearby} {k} {j}")
        np.testing.assert_array_equal(r["flat"][j].view(np.int32), r["thread"][j].view(np.int32), err_msg=f"flat vs thread {nearby} {k} {j}")
    o.set_nearby(nearby)
    sub = rng.integers(0, nq, min(nq, 3000))
    oi, od, _, oc = o.knn(q[sub], k, 5.0)
    np.testing.assert_array_equal(r["flat"][0][sub], oi)
    np.testing.assert_array_equal(r["flat"][1][sub].view(np.int32), od.view(np.int32))
    np.testing.assert_array_equal(r["flat"][2][sub], oc)
    print("batch ok", nearby, k, nq, "found-all", float((r["flat"][2] == k).mean()))
o.set_nearby(18)

# 2. 40-point buckets: overflow levels inside phase B, candidate lists that overflow (the serial re-walk), entry lists
#    that fill up within one pass (NEARBY74 on a dense block)
gm = lsdreg.HashVoxelMap(0.5, 14)
pts = np.zeros((4000, 4), np.float32)
pts[:, :3] = rng.uniform(-0.2, 0.2, (4000, 3)) + rng.integers(0, 10, (4000, 1)) * np.array([[0.5, 0, 0]])
gm.insert(pts, 0)
dense = np.zeros((6000, 4), np.float32)
dense[:, :3] = rng.uniform(-1.6, 1.6, (6000, 3)) + np.array([40.0, 0.0, 0.0])      # every cell of a 7^3 block holds ~17 points
gm.insert(dense, 4000)
qq = np.concatenate([np.repeat(pts[:300], 20, axis=0), dense[:3000]]).copy()
for nearby, k in ((18, 5), (74, 5), (18, 1), (26, 5)):
    r = shapes(gm, qq, k, nearby)
    for j in range(3):
        np.testing.assert_array_equal(r["flat"][j].view(np.int32), r["warp"][j].view(np.int32), err_msg=f"dense flat vs warp {nearby} {k} {j}")
    print("dense ok", nearby, k)

# 3. inside the LIO scan: same Nearest_Points, same posterior, with either search shape
Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
tgt = synth.block_center(0, 0) + np.array([1.0, -2.0, 0.0])
scan = synth.scan64(2, 250, Rgt, tgt)
dR, dt = synth.perturb(5)
prior = eskf.State(); prior.rot = eskf.R_to_quat(Rgt @ dR); prior.pos = tgt + dt
res = []
for shape in (0, 3, 4):
    f = lsdreg.LioFrontend(map_log2_lines=20)
    f.map.insert(m, 0); f.set_next_id(m.shape[0])
    f.set_knn_shape(shape)
    for nearby in (lsdreg.STENCIL_NEARBY18, lsdreg.STENCIL_NEARBY74):
        f.set_nearby(nearby)
        x, P, info = f.scan(scan, prior.to_vec(), lsdreg.init_cov())
        mt = f.get_matches()
        res.append((shape, nearby, x.copy(), P.copy(), mt["idx"].copy(), info["n_eff"], f.map.stats()))
for a, b in ((res[0], res[2]), (res[1], res[3])):
    assert a[1] == b[1]
    np.testing.assert_array_equal(a[4], b[4])
    np.testing.assert_array_equal(a[2], b[2]); np.testing.assert_array_equal(a[3], b[3])
    assert a[5] == b[5] and a[6] == b[6]
    print("lio ok", a[1], "n_eff", a[5])
# shape 4 = the flat search fused with the plane fit and the reduction: same Nearest_Points, same rows, another reduction
# tree (64 points per block instead of 256) -> sums and pose equal to rounding
for a, b in ((res[0], res[4]), (res[1], res[5])):
    assert a[1] == b[1]
    np.testing.assert_array_equal(a[4], b[4])
    assert a[5] == b[5] and a[6] == b[6]
    np.testing.assert_allclose(a[2], b[2], rtol=0, atol=1e-11); np.testing.assert_allclose(a[3], b[3], rtol=1e-5, atol=1e-13)
    print("lio fused ok", a[1], "n_eff", a[5], "max |dx|", float(np.abs(a[2] - b[2]).max()))
print("FLAT_OK")
'''


@pytest.mark.xfail(strict=False, reason="never run on a GPU yet (written after the round's GPU budget was spent); see the module docstring")
def test_flat_shape_is_bit_identical_to_the_validated_shapes():
    r = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": _ROOT}], cwd=_ROOT, capture_output=True, text=True, timeout=420)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0 and "FLAT_OK" in r.stdout, tail
