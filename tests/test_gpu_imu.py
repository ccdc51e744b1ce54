"""GPU parity, row N1: ImuProcess (IMU initialisation, forward propagation on the host, per-point undistortion on the
device) against the numpy restatement, through the C ABI.  Bars: state / covariance / IMU pose list to 1e-11; undistorted
points equal to float32 rounding (both sides evaluate IMU_Processing.hpp:390 in double and store floats)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stream(n_frames, rng, n_pts=20000, t_min_ms=0.0, omega=(0.0, 0.0, 0.3), accel=(0.5, 0.0, 0.0), imu_hz=200, ext=None):
    """Measurement groups of a platform that stands still for the first 11 frames (IMU init needs > 100 samples) and then
    turns / accelerates.  IMU: gyr rad/s, acc in g-units (reads +1 g on z at rest)."""
    frames = []
    per = imu_hz // 10
    for f in range(n_frames):
        beg = 0.1 * f
        moving = f >= 11
        stamps = beg + (np.arange(per) + 1) * (0.1 / per)
        imu = np.zeros((per, 7))
        imu[:, 0] = stamps
        imu[:, 1:4] = rng.normal(0, 0.002, (per, 3)) + (np.array(omega) if moving else 0.0)
        imu[:, 4:7] = np.array([0.0, 0.0, 1.0]) + rng.normal(0, 0.002, (per, 3)) + (np.array(accel) / 9.81 if moving else 0.0)
        pts = np.zeros((n_pts, 4), np.float32)
        pts[:, :3] = rng.uniform(-40, 40, (n_pts, 3)); pts[:, 2] *= 0.1; pts[:, 3] = rng.uniform(0, 255, n_pts)
        tms = rng.uniform(t_min_ms, 100.0, n_pts).astype(np.float32)     # NOT sorted: the library must not need sorted input
        if t_min_ms == 0.0:
            tms[rng.integers(0, n_pts, 5)] = 0.0
        frames.append(dict(lidar_beg_time=beg, lidar_end_time=beg + 0.1, points=pts, time_ms=tms, imu=imu, ins_vel=None))
    return frames


@pytest.mark.parametrize("t_min_ms,ext", [(0.0, False), (0.7, True)])
def test_imu_process_matches_oracle(t_min_ms, ext):
    import lsdreg
    from lsdreg import synth
    from oracle import eskf as E
    from oracle.imu import OracleImuProcess
    rng = np.random.default_rng(5)
    ext_R = synth.rot_from_rpy(0.02, -0.01, 0.05) if ext else np.eye(3)
    ext_t = np.array([0.3, -0.1, 0.2]) if ext else np.zeros(3)
    frames = _stream(15, rng, t_min_ms=t_min_ms)
    g = lsdreg.ImuProcess(ext_R=ext_R, ext_t=ext_t)
    o = OracleImuProcess(ext_R=ext_R, ext_t=ext_t)
    xo, Po = E.State(), E.init_P()
    xg, Pg = xo.to_vec(), Po.copy()
    n_und = 0
    for f, m in enumerate(frames):
        out_o = o.process(m, xo, Po)
        st, xg, Pg, n = g.process(m["imu"], m["lidar_beg_time"], m["lidar_end_time"], m["points"], m["time_ms"], xg, Pg)
        np.testing.assert_allclose(xg, xo.to_vec(), rtol=0, atol=1e-11)
        np.testing.assert_allclose(Pg, Po, rtol=1e-10, atol=1e-14)
        if out_o is None:
            assert st == lsdreg.IMU_INITIALIZING and n == 0 and not g.is_init()
            continue
        assert st == lsdreg.OK and n == len(m["points"]) and g.is_init()
        n_und += 1
        poses = g.poses()
        assert len(poses) == len(o.IMUpose)
        for pg, po in zip(poses, o.IMUpose):
            np.testing.assert_allclose(pg[0], po["t"], atol=1e-12)
            np.testing.assert_allclose(pg[1:4], po["acc"], atol=1e-10); np.testing.assert_allclose(pg[4:7], po["gyr"], atol=1e-12)
            np.testing.assert_allclose(pg[7:10], po["vel"], atol=1e-11); np.testing.assert_allclose(pg[10:13], po["pos"], atol=1e-11)
            np.testing.assert_allclose(pg[13:22].reshape(3, 3), po["rot"], atol=1e-12)
        order = np.argsort(m["time_ms"], kind="stable")          # the oracle returns the cloud sorted by time
        cg = g.cloud()[order]
        np.testing.assert_array_equal(cg[:, 3], out_o[:, 3])
        d = np.abs(cg[:, :3].astype(np.float64) - out_o[:, :3])
        assert d.max() <= 8e-6, d.max()                          # <= 1-2 float32 ulp at 40 m
        assert (d == 0).mean() > 0.99                            # and bit-identical almost everywhere
        moved = np.abs(cg[:, :3] - m["points"][order][:, :3]).max()
        assert moved > (0.05 if f >= 11 else 0.0)                # undistortion does something once the platform moves
    assert n_und >= 4


def test_undistort_disabled_passes_points_through():
    import lsdreg
    from oracle import eskf as E
    rng = np.random.default_rng(6)
    frames = _stream(13, rng, n_pts=5000)
    g = lsdreg.ImuProcess(undistort=0)
    xg, Pg = E.State().to_vec(), E.init_P()
    for m in frames:
        st, xg, Pg, n = g.process(m["imu"], m["lidar_beg_time"], m["lidar_end_time"], m["points"], m["time_ms"], xg, Pg)
    assert st == lsdreg.OK and n == 5000
    np.testing.assert_array_equal(g.cloud(), frames[-1]["points"])


def test_imu_then_lio_pipeline_tracks_a_moving_sensor():
    """Process (IMU) -> lsd_lio_scan on the undistorted cloud, the order of fastlio_main (laserMapping.cpp:1188-1292):
    the pose follows a sensor moving at constant velocity through the block scene, and undistortion is what makes the
    motion-distorted scans consistent (the same run with undistort=0 is measurably worse)."""
    import torch
    import lsdreg
    from lsdreg import synth
    from oracle import eskf as E

    def run(undistort):
        rng = np.random.default_rng(9)
        m0 = synth.block_map(3, 1, 1, 0.5)
        lio = lsdreg.LioFrontend(map_log2_lines=20)
        lio.map.insert(m0, 0)
        lio.set_next_id(m0.shape[0])
        imu = lsdreg.ImuProcess(undistort=undistort)
        c = synth.block_center(0, 0)
        x = lsdreg.make_state(pos=c)
        P = E.init_P()
        v = np.array([6.0, 2.0, 0.0])                               # m/s once moving; the IMU only ever reads gravity
        per, pos, errs = 20, c.copy(), []
        for f in range(17):
            beg = 0.1 * f
            stamps = beg + (np.arange(per) + 1) * (0.1 / per)
            im = np.zeros((per, 7)); im[:, 0] = stamps; im[:, 4:7] = [0, 0, 1.0]
            im[:, 1:4] = rng.normal(0, 1e-4, (per, 3)); im[:, 4:7] += rng.normal(0, 1e-4, (per, 3))
            moving = f >= 11
            if moving:
                pos = pos + v * 0.1                                 # sensor position at the END of this scan
            scan = synth.scan64(100 + f, 400, np.eye(3), pos)
            tms = np.linspace(0, 100, scan.shape[0]).astype(np.float32)
            if moving:                                              # a point fired at t was seen from pos_end - v (T - t)
                scan[:, :3] += (v[None, :] * (0.1 - tms[:, None] / 1000.0)).astype(np.float32)
            if f == 11:
                x[14:17] = v                                        # the reference seeds velocity from INS (IMU_Processing.hpp:198-201)
            st, x, P, n = imu.process(im, beg, beg + 0.1, scan, tms, x, P)
            if st == lsdreg.IMU_INITIALIZING:
                continue
            x, P, info = lio.scan(torch.from_numpy(imu.cloud()).cuda(), x, P)
            if moving:
                errs.append(float(np.abs(x[0:3] - pos).max()))
        return errs

    e_on, e_off = run(1), run(0)
    assert len(e_on) >= 5 and max(e_on[-3:]) < 0.05, e_on
    assert np.mean(e_off[-3:]) > 2 * np.mean(e_on[-3:]), (e_on, e_off)
