"""GPU parity, row N2: local-map assembly (key-frame selection + concatenation + voxel grid) against the numpy
restatement, and the hand-over of the device-resident result to the matcher."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _keyframes():
    from lsdreg import synth
    frames = []
    c = synth.block_center(0, 0)
    for k in range(24):                                           # a drive through the block, one key frame per 2.5 m
        pos = c + np.array([-30.0 + 2.5 * k, 4.0 * np.sin(0.3 * k), 0.0])
        R = synth.rot_from_rpy(0, 0, 0.05 * k)
        s = synth.scan64(200 + k, 300, R, pos)
        w = s.copy()
        w[:, :3] = (s[:, :3].astype(np.float64) @ R.T + pos).astype(np.float32)   # mTransfromPoints: map frame
        frames.append((w, pos))
    return frames


def test_local_map_matches_oracle_and_feeds_the_matcher():
    import lsdreg
    from lsdreg import synth
    from oracle.localmap import OracleLocalMap
    frames = _keyframes()
    g = lsdreg.LocalMap(0.5, 1.0)
    o = OracleLocalMap(0.5, 1.0)
    for w, pos in frames:
        g.add_keyframe(w, pos); o.add_keyframe(w, pos)
    c = synth.block_center(0, 0)
    for pose in (c + [-10.0, 1.0, 0.0], c + [20.0, -3.0, 0.0], frames[0][1] + [-8.0, 0.0, 0.0]):
        want, nk = o.update(pose)
        st, n, k, dist = g.update(pose)
        assert st == lsdreg.OK and k == nk and want is not None
        got = g.cloud()
        assert got.shape == want.shape and n == len(want) and n > 1000
        np.testing.assert_allclose(got[:, :3], want[:, :3], rtol=4e-6, atol=2e-5)   # PCL sums float32 in order (dozens of points per voxel at |x| ~ 100 m); the device sums exactly
        np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=1e-5, atol=4e-3)
    # nearest key frame 20-30 m away: found by the radius search but rejected; > 30 m: out of map
    for off, in_radius in ((25.0, True), (45.0, False)):
        pose = frames[0][1] + [-off, 0.0, 0.0]
        want, nk = o.update(pose)
        st, n, k, dist = g.update(pose)
        assert want is None and st == lsdreg.LOCALMAP_NONE and n == 0 and (k > 0) == in_radius and len(g.cloud()) == 0
        assert abs(dist - off) < 1.0
    # updateLocalMap(mLocalMap): the assembled map goes to the matcher without leaving the device
    pose = c + [5.0, 2.0, 0.0]
    st, n, k, dist = g.update(pose)
    ptr, n_dev = g.cloud_dev()
    assert st == lsdreg.OK and n_dev == n
    R = synth.rot_from_rpy(0.0, 0.0, 0.4)
    scan = synth.scan64(999, 300, R, pose)
    m = lsdreg.Matcher("FAST_VGICP")
    m.set_target_ptr(ptr, n_dev)
    m.set_source(scan)
    dR, dt = synth.perturb(3, 0.3, 1.5)
    guess = np.eye(4); guess[:3, :3] = R @ dR; guess[:3, 3] = pose + dt
    T = m.align(guess)
    assert m.converged and np.abs(T[:3, 3] - pose).max() < 0.1


def test_local_map_caps_at_200k_points():
    import lsdreg
    from oracle.localmap import OracleLocalMap
    rng = np.random.default_rng(2)
    g = lsdreg.LocalMap(0.3, 0.0)
    o = OracleLocalMap(0.3, 0.0)
    for k in range(12):
        pts = np.zeros((30000, 4), np.float32)
        pts[:, :3] = rng.uniform(-25, 25, (30000, 3)); pts[:, 2] *= 0.1
        pos = np.array([0.5 * k, 0.0, 0.0])
        g.add_keyframe(pts, pos); o.add_keyframe(pts, pos)
    want, nk = o.update([0.0, 0.0, 0.0])
    st, n, k, dist = g.update([0.0, 0.0, 0.0])
    assert st == lsdreg.OK and k == nk == 12 and n == len(want)           # 7 key frames reach 210 000 >= 200 000 points
    np.testing.assert_allclose(g.cloud()[:, :3], want[:, :3], rtol=4e-6, atol=2e-5)
