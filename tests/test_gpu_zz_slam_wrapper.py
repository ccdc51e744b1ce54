"""GPU: the per-frame entries of the pybind `slam_wrapper` module (init_slam / setup_slam / process / update_odom,
reference slam/src/slam_wrapper.cpp:6-18,62-135,194-224) driven exactly as slam/slam.py drives them (slam.py:49-87,234-244:
argument order, dict layouts, numpy dtypes and shapes), against the same sensor stream fed to the C++ seam directly
(lsdreg.FastLio = lsd_fastlio_*, itself held to the compiled reference by tests/test_gpu_zz_fastlio_seam.py).
Bars: odom_matrix equals the seam's odometry (conjugated with the IMU-INS extrinsic) to 1e-9; heading / pitch / roll follow
SLAM::run's convention (slam.cpp:349-363); key frames come out filtered and posed.  Subprocess: the module keeps one
process-global SLAM object, like the reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests"); sys.path.insert(0, %(root)r + "/lidar-slam-detection_b200")
import numpy as np
import lsdreg
import slam_wrapper as slam
import test_oracle_fastlio as T

lsdreg.init(0)
sensors = slam.init_slam("mapping", "/tmp/map", "FastLIO", ["0-Ouster-OS1-64", "1-VLP-16", "RTK", "IMU", "front_camera"], 0.2, 4.0, 20.0, 50.0)
assert sensors == ["RTK", "IMU", "0-Ouster-OS1-64", "1-VLP-16", "front_camera"], sensors       # HDL_FastLIO::setSensors order
slam.set_ins_external_param(0.3, -0.1, 0.2, 2.0, 1.0, -1.5)       # x y z yaw pitch roll (deg), slam.py:59-60 passes [5],[4],[3]
slam.set_imu_external_param(0.1, 0.05, 0.0, 0.0, 0.0, 0.0)
assert slam.setup_slam() is True

def rpyt(x, y, z, yaw, pitch, roll):
    d = np.pi / 180
    cy, sy, cp, sp, cr, sr = np.cos(yaw * d), np.sin(yaw * d), np.cos(pitch * d), np.sin(pitch * d), np.cos(roll * d), np.sin(roll * d)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]); Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]]); Ry = np.array([[cr, 0, sr], [0, 1, 0], [-sr, 0, cr]])
    M = np.eye(4); M[:3, :3] = Rz @ Rx @ Ry; M[:3, 3] = [x, y, z]
    return M
Ts, Ti = rpyt(0.3, -0.1, 0.2, 2.0, 1.0, -1.5), rpyt(0.1, 0.05, 0.0, 0, 0, 0)
M = Ti @ np.linalg.inv(Ts)                                          # mImuInsStaticTrans (fastlio.cpp:156)
g = lsdreg.FastLio(M[:3, :3], M[:3, 3], map_log2_lines=18)          # the seam, fed the same things
n_pose, n_key = 0, 0
for f, (imu, raw, stamp_us, hdr) in enumerate(T._stream(14, np.eye(3), np.zeros(3))):
    # slam.py's input_dict: imu rows (t us, gyr deg/s, acc g) f64; points {name: f32 [N,4]}; points_attr {name: {points_attr f32 [N,2], timestamp}}
    imu_list = np.concatenate([imu[:, :1] * 1e6, imu[:, 1:4] * 180.0 / np.pi, imu[:, 4:7]], 1)
    attr = np.stack([stamp_us.astype(np.float32), np.zeros(raw.shape[0], np.float32)], 1)
    res = slam.process({"0-Ouster-OS1-64": raw, "1-VLP-16": raw[:10]}, {"0-Ouster-OS1-64": {"points_attr": attr, "timestamp": hdr}, "1-VLP-16": {"points_attr": attr[:10], "timestamp": hdr}},
                       {}, {}, {}, {"latitude": 31.0, "longitude": 121.0, "altitude": 4.0, "Status": 1, "Sensor": "GNSS"}, imu_list, hdr)
    assert set(res) == {"frame_start_timestamp", "pose", "slam_valid"} and res["slam_valid"] is True and res["frame_start_timestamp"] == hdr
    pose = res["pose"]
    assert set(pose) == {"latitude", "longitude", "altitude", "heading", "pitch", "roll", "Ve", "Vn", "Vu", "Status", "state", "timestamp", "odom_matrix"}
    assert pose["state"] == "Mapping" and pose["latitude"] == 31.0 and pose["Status"] == 1 and pose["timestamp"] == hdr
    Tm = pose["odom_matrix"]
    assert Tm.shape == (4, 4) and Tm.dtype == np.float64
    # the same frame through the seam: transform to the INS frame, enqueue, main, odometry, conjugate
    for row in imu:
        g.push_imu(row[0], row[1:4], row[4:7])
    pts = raw.copy(); pts[:, :3] = (raw[:, :3].astype(np.float64) @ Ts[:3, :3].T + Ts[:3, 3]).astype(np.float32)
    g.push_scan(pts, stamp_us, hdr)
    stepped = g.step()
    if stepped:
        s16, e16 = g.odometry()
        want = np.linalg.inv(M) @ s16 @ M
        # (the wrapper moves the points to the INS frame in its own arithmetic: a few inputs differ from `pts` by one fp32
        # ulp, which reaches the pose at the 1e-9 level — 1.1e-9 measured)
        np.testing.assert_allclose(Tm, want, rtol=0, atol=1e-8)
        n_pose += 1
        R = Tm[:3, :3]
        h = (-np.degrees(np.arctan2(-R[0, 1], R[1, 1]))) %% 360.0     # small roll / pitch: heading = -yaw of eulerAngles(2, 0, 1), in [0, 360)
        assert abs(((pose["heading"] - h + 180) %% 360) - 180) < 1e-6, (pose["heading"], h)
        assert abs(pose["pitch"]) < 5 and abs(pose["roll"]) < 5
    else:
        np.testing.assert_array_equal(Tm, np.eye(4))
    out = slam.update_odom()
    assert set(out) == {"odoms", "keyframes"}
    for kf in out["keyframes"]:
        assert set(kf) == {"points", "image", "pose", "stamp"} and kf["points"].dtype == np.float32 and kf["points"].shape[1] == 4
        assert 0 < kf["points"].shape[0] <= raw.shape[0] and kf["pose"].shape == (4, 4) and kf["stamp"] == hdr
        assert np.abs(kf["points"][:, :2]).max() < 50.0               # pointsDistanceFilter(0, key_frames_range)
        n_key += 1
    assert len(out["odoms"]) == n_key and all(v.shape == (4, 4) for v in out["odoms"].values())
assert n_pose >= 10 and n_key >= 1, (n_pose, n_key)
slam.deinit_slam()
try:
    slam.init_slam("mapping", "/tmp/map", "RTKM", ["RTK"], 0.2, 4.0, 20.0, 50.0)
    raise SystemExit("RTKM must be refused")
except ValueError:
    pass
print("SLAM_WRAPPER_OK", n_pose, n_key)
'''


@pytest.mark.gpu
def test_process_as_slam_py_calls_it():
    r = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": _ROOT}], cwd=_ROOT, capture_output=True, text=True, timeout=420)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0 and "SLAM_WRAPPER_OK" in r.stdout, tail


def test_init_slam_sensor_rules_without_a_device():
    """HDL_FastLIO::setSensors (fastlio.cpp:118-151) as init_slam returns it; needs no GPU."""
    sys.path.insert(0, os.path.join(_ROOT, "lidar-slam-detection_b200"))
    import slam_wrapper as slam
    assert slam.init_slam("mapping", "", "FastLIO", ["0-A", "RTK"], 0.2, 4.0, 20.0, 50.0) == ["RTK"]            # no IMU: lidars dropped
    assert slam.init_slam("mapping", "", "FastLIO", ["cam", "1-B", "IMU", "0-A"], 0.2, 4.0, 20.0, 50.0) == ["IMU", "cam", "1-B", "0-A"]
    with pytest.raises(ValueError):
        slam.init_slam("localization", "", "FastLIO", ["IMU", "0-A"], 0.2, 4.0, 20.0, 50.0)
    slam.init_slam("mapping", "", "FastLIO", ["RTK"], 0.2, 4.0, 20.0, 50.0)
    with pytest.raises(RuntimeError):
        slam.setup_slam()                                    # no lidar registered
    slam.deinit_slam()
    with pytest.raises(RuntimeError):
        slam.update_odom()
