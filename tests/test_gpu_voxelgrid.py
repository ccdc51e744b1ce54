"""GPU parity: voxel-grid downsample (K1) against the restated PCL VoxelGrid.
Bar: identical leaf set and output order; centroids within 1e-5 m (PCL sums fp32 in sort order,
we sum exactly in fixed point — SURVEY.md §7 "PCL VoxelGrid centroid summation order")."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(scan, leaf):
    import lsdreg
    from oracle import oracle as O
    ref, vidx = O.voxelgrid(scan, leaf, want_vidx=True)
    vg = lsdreg.VoxelGrid(max(scan.shape[0], 1))
    out = vg.filter(scan, leaf)
    assert out.shape == ref.shape
    # every bit of every channel: the centroids are the restated pcl::VoxelGrid's sequential fp32 sums in input order
    # (LSD_VG_SUMS=fixed selects the exact fixed-point sums of round 1, <= 1e-5 m / 2e-3 intensity units from these)
    if os.environ.get("LSD_VG_SUMS", "")[:1] == "f":
        np.testing.assert_allclose(out[:, :3], ref[:, :3], rtol=0, atol=1e-5)
        np.testing.assert_allclose(out[:, 3], ref[:, 3], rtol=0, atol=2e-3)  # intensity 0..255
    else:
        assert (out.view(np.uint32) == ref.view(np.uint32)).all()
    out2 = vg.filter(scan, leaf)  # scratch returned to zero; bit-stable run to run
    assert (out2.view(np.int32) == out.view(np.int32)).all()
    return out


def test_voxelgrid_scan_16k(small_world):
    _check(small_world["scan"], 0.5)


def test_voxelgrid_scan_100k_and_leaf_02():
    from lsdreg import synth
    scan = synth.scan64(3, 1563)
    out = _check(scan, 0.5)
    assert 10000 < out.shape[0] <= 100000  # reference static cap, laserMapping.cpp:86
    _check(scan[::3], 0.2)


def test_voxelgrid_edge_cases():
    import lsdreg
    vg = lsdreg.VoxelGrid(1000)
    assert vg.filter(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)
    one = np.array([[1.0, -2.0, 3.0, 7.0]], np.float32)
    np.testing.assert_array_equal(vg.filter(one, 0.5), one)
    dup = np.repeat(one, 100, 0)
    np.testing.assert_allclose(vg.filter(dup, 0.5), one, atol=1e-6)
    # grid above INT32_MAX leaves: PCL warns and returns the input unchanged
    far = np.array([[0, 0, 0, 1], [5e5, 5e5, 4e3, 2]], np.float32)
    np.testing.assert_array_equal(vg.filter(far, 0.5), far)


def test_non_finite_points_are_skipped():
    """One Inf / NaN coordinate must not blow the grid up (it used to make the leaf indices wild: out-of-bounds bitmap
    writes).  pcl::VoxelGrid skips non-finite points; the finite ones give exactly what they give alone."""
    import lsdreg
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    fin = np.zeros((30000, 4), np.float32)
    fin[:, :3] = rng.uniform(-40, 40, (30000, 3)) * [1, 1, 0.1]
    fin[:, 3] = rng.uniform(0, 255, 30000)
    bad = np.array([[np.inf, 0, 0, 1], [0, -np.inf, 0, 1], [np.nan, 1, 1, 1], [1, 1, np.nan, 1]], np.float32)
    mixed = np.concatenate([fin[:100], bad[:2], fin[100:20000], bad[2:], fin[20000:]])
    vg = lsdreg.VoxelGrid(max_points=40000)
    want = O.voxelgrid(fin, 0.5)
    for _ in range(2):                     # twice: the scratch must come back clean
        got = vg.filter(mixed, 0.5)
        assert got.shape == want.shape
        np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=0, atol=1e-5)
        np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=0, atol=2e-3)
    assert vg.filter(bad, 0.5).shape[0] == 0
    got2 = vg.filter(fin, 0.5)
    assert (got2.view(np.int32) == got.view(np.int32)).all()      # bit-identical to the run that had to skip points


def test_crowded_leaves_sum_in_input_order():
    """Leaves with 33 ... 700 points (the warp-cooperative path of vg_sum_kernel) beside sparse ones, points of a leaf scattered
    all over the input, values spanning five orders of magnitude so that every change of summation order shows: every bit of
    every centroid equals the restated pcl::VoxelGrid's (sequential fp32 sums in input order)."""
    rng = np.random.default_rng(12)
    parts = [rng.uniform(-30, 30, (20000, 4)).astype(np.float32)]
    for k in (33, 64, 65, 200, 700):
        c = np.floor(rng.uniform(-20, 20, 3) / 0.5) * 0.5
        pts = (c + rng.uniform(0.01, 0.49, (k, 3))).astype(np.float32)
        parts.append(np.concatenate([pts, (10.0 ** rng.uniform(-2, 3, (k, 1))).astype(np.float32)], 1))
    scan = np.concatenate(parts)
    scan = np.ascontiguousarray(scan[rng.permutation(scan.shape[0])])
    _check(scan, 0.5)
