"""CPU checks of the IMU restatement (oracle/imu.py, oracle/eskf.py::predict; row N1).  IKFoM needs Boost, so ImuProcess
itself cannot be compiled here (parity unpinned); what CAN be pinned is pinned — so3_math.h::Exp against the compiled
reference header — and the rest is held to the physics it implements."""
import numpy as np
import pytest

from oracle import eskf as E
from oracle import oracle as O
from oracle.imu import OracleImuProcess, exp_rodrigues


@pytest.mark.skipif(not (O.HAVE_REF and hasattr(O.ref, "ref_so3_exp")), reason="oracle/_ref/libref_lio.so without ref_so3_exp")
def test_exp_matches_compiled_reference():
    rng = np.random.default_rng(0)
    for scale in (1e-9, 1e-3, 0.5, 3.0):
        for _ in range(5):
            w = rng.normal(0, scale, 3); dt = float(rng.uniform(0, 0.1))
            out = np.zeros(9)
            O.ref.ref_so3_exp(np.ascontiguousarray(w), dt, out)
            np.testing.assert_allclose(exp_rodrigues(w, dt), out.reshape(3, 3), rtol=0, atol=1e-15)


def _frame(beg, n_pts, rng, gyr=(0, 0, 0), acc=(0, 0, 1.0), per=20):
    stamps = beg + (np.arange(per) + 1) * (0.1 / per)
    imu = np.zeros((per, 7)); imu[:, 0] = stamps; imu[:, 1:4] = gyr; imu[:, 4:7] = acc
    pts = np.zeros((n_pts, 4), np.float32); pts[:, :3] = rng.uniform(-30, 30, (n_pts, 3))
    tms = np.sort(rng.uniform(0, 100, n_pts)).astype(np.float32)
    return dict(lidar_beg_time=beg, lidar_end_time=beg + 0.1, points=pts, time_ms=tms, imu=imu, ins_vel=None)


def _initialised(rng, gyr=(0, 0, 0)):
    imu = OracleImuProcess()
    x, P = E.State(), E.init_P()
    f = 0
    while True:
        out = imu.process(_frame(0.1 * f, 100, rng), x, P)
        f += 1
        if not imu.imu_need_init:
            return imu, x, P, f


def test_initialisation_takes_more_than_100_samples_and_finds_gravity():
    rng = np.random.default_rng(1)
    imu, x, P, f = _initialised(rng)
    assert f == 5 and imu.init_iter_num == 101                     # 20 samples per frame: N = 101 > MAX_INI_COUNT after the 5th
    np.testing.assert_allclose(x.grav, [0, 0, -E.S2_LEN], atol=1e-9)   # IMU reads +1 g on z at rest -> gravity points down
    np.testing.assert_allclose(P, E.init_P())


def test_static_platform_is_left_alone_and_rotation_is_compensated():
    rng = np.random.default_rng(2)
    imu, x, P, f = _initialised(rng)
    m = _frame(0.1 * f, 2000, rng)
    out = imu.process(m, x, P)
    np.testing.assert_allclose(out[:, :3], m["points"][:, :3], atol=2e-5)      # nothing moves: undistortion is the identity
    # the reference scales the accelerometer to 9.81 m/s^2 but keeps gravity on a sphere of 9.809: 1 mm/s^2 of phantom lift
    np.testing.assert_allclose(x.pos, 0, atol=1e-5); np.testing.assert_allclose(x.vel, 0, atol=1.1e-4)
    assert np.linalg.eigvalsh(0.5 * (P + P.T)).min() > 0
    # constant yaw rate: a point seen at time t is rotated by the yaw still to come, w * (T - t)
    w = 0.5
    m = _frame(0.1 * (f + 1), 2000, rng, gyr=(0, 0, w))
    out = imu.process(m, x, P)
    order = np.argsort(m["time_ms"], kind="stable")
    raw = m["points"][order, :3].astype(np.float64); t = m["time_ms"][order].astype(np.float64) / 1000.0
    ang = -w * (0.1 - t)                                                       # into the end frame: rotate back by the remaining yaw
    want = np.stack([np.cos(ang) * raw[:, 0] - np.sin(ang) * raw[:, 1], np.sin(ang) * raw[:, 0] + np.cos(ang) * raw[:, 1], raw[:, 2]], 1)
    sel = t > 0.006                                                            # the first IMU segment starts from the (static) previous rate
    assert np.abs(out[sel, :3] - want[sel]).max() < 2e-3                       # piecewise-constant rate vs the exact arc: sub-millimetre here
    R_end = E.quat_to_R(x.rot)
    # the filter turned by w * 0.1 s minus half of the first 5 ms interval (midpoint of the previous, static, sample and the first moving one)
    assert abs(np.arctan2(R_end[1, 0], R_end[0, 0]) - (w * 0.1 - 0.5 * w * 0.005)) < 1e-6


def test_predict_keeps_the_covariance_symmetric_and_grows_it():
    rng = np.random.default_rng(3)
    x = E.State(); x.boxplus(rng.normal(0, 0.2, 23))
    P = E.init_P()
    Q = np.diag([0.1] * 6 + [1e-4] * 6)
    tr0 = np.trace(P)
    for _ in range(20):
        P = E.predict(x, P, 0.005, Q, np.array([0.1, 0.0, 9.8]), np.array([0.01, -0.02, 0.3]))
    np.testing.assert_allclose(P, P.T, atol=1e-12)
    assert np.trace(P) > tr0 and np.linalg.eigvalsh(P).min() > 0
    np.testing.assert_allclose(np.linalg.norm(x.grav), E.S2_LEN, rtol=1e-12)   # gravity stays on its sphere
    np.testing.assert_allclose(np.linalg.norm(x.rot), 1.0, rtol=1e-12)
