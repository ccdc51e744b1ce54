"""TEST INFRASTRUCTURE ONLY — numpy restatement of ImuProcess (row N1):
IMU initialisation, forward propagation and per-point undistortion of
/root/reference/slam/mapping/fastlio/src/IMU_Processing.hpp:167-450 on top of oracle/eskf.py::predict.

Pin status: PINNED.  The reference's ImuProcess compiles unmodified with the Boost.Preprocessor / PCL shims of
oracle/ref_shim_ikfom (oracle/ref_ikfom.cpp -> oracle/_ref/libref_ikfom.so); tests/test_oracle_ikfom.py runs whole IMU
streams through both: state 1e-11, covariance 1e-9, the IMU pose list, and every undistorted point (>= 99 % bit-identical,
the rest within 1 float32 ulp), including the reference's repeated compensation of the earliest point.
Written independently of the product's C++ (csrc/imu.cu).

A measurement group is (lidar_beg_time, lidar_end_time, points [n,4] float32 (x, y, z, intensity),
time_ms [n] float32 (PointType::curvature), imu [m,7] float64 (stamp s, gyr xyz rad/s, acc xyz in g-units),
ins_vel (3,) or None).
"""
from __future__ import annotations

import numpy as np

from . import eskf as E

G_M_S2 = 9.81          # common_lib.h:21
MAX_INI_COUNT = 100    # IMU_Processing.hpp:26


def exp_rodrigues(ang_vel, dt):
    """so3_math.h:36-58"""
    n = np.linalg.norm(ang_vel)
    if n > 0.0000001:
        K = E.hat(ang_vel / n)
        a = n * dt
        return np.eye(3) + np.sin(a) * K + (1.0 - np.cos(a)) * (K @ K)
    return np.eye(3)


class OracleImuProcess:
    def __init__(self, ext_R=None, ext_t=None, gyr_cov=0.1, acc_cov=0.1, b_gyr_cov=0.0001, b_acc_cov=0.0001, undistort=True):
        self.Lidar_R = np.eye(3) if ext_R is None else np.asarray(ext_R, np.float64).reshape(3, 3)
        self.Lidar_T = np.zeros(3) if ext_t is None else np.asarray(ext_t, np.float64)
        self.cov_gyr_scale = np.full(3, gyr_cov); self.cov_acc_scale = np.full(3, acc_cov)
        self.cov_bias_gyr = np.full(3, b_gyr_cov); self.cov_bias_acc = np.full(3, b_acc_cov)
        self.undistort = undistort
        self.Q = np.diag([1e-4] * 6 + [1e-5] * 6)       # process_noise_cov(), use-ikfom.hpp:36-44
        self.b_first_frame = True
        self.imu_need_init = True
        self.state_init_done = False
        self.first_lidar_time = 0.0
        self._reset()

    def _reset(self):                                    # :107-120
        self.cov_acc = np.full(3, 0.1); self.cov_gyr = np.full(3, 0.1)
        self.mean_acc = np.array([0.0, 0.0, -1.0]); self.mean_gyr = np.zeros(3)
        self.vel_last = np.zeros(3); self.angvel_last = np.zeros(3); self.acc_s_last = np.zeros(3)
        self.imu_need_init = True; self.state_init_done = False
        self.init_iter_num = 1
        self.last_imu = np.zeros(7)
        self.last_lidar_end_time = 0.0
        self.IMUpose = []

    # ---------------------------------------------------------------- IMU_init :167-235
    def _imu_init(self, meas, x: E.State, P: np.ndarray):
        imu = meas["imu"]
        if self.b_first_frame:
            self._reset()
            self.init_iter_num = 1
            self.b_first_frame = False
            self.mean_acc = imu[0, 4:7].copy(); self.mean_gyr = imu[0, 1:4].copy()
            self.first_lidar_time = meas["lidar_beg_time"]
        N = self.init_iter_num
        for s in imu:
            cur_acc, cur_gyr = s[4:7], s[1:4]
            self.mean_acc = self.mean_acc + (cur_acc - self.mean_acc) / N
            self.mean_gyr = self.mean_gyr + (cur_gyr - self.mean_gyr) / N
            self.cov_acc = self.cov_acc * (N - 1.0) / N + (cur_acc - self.mean_acc) * (cur_acc - self.mean_acc) * (N - 1.0) / (N * N)
            self.cov_gyr = self.cov_gyr * (N - 1.0) / N + (cur_gyr - self.mean_gyr) * (cur_gyr - self.mean_gyr) * (N - 1.0) / (N * N)
            N += 1
        self.init_iter_num = N
        if meas.get("ins_vel") is not None:
            self.vel_last = np.asarray(meas["ins_vel"], np.float64)
        na = np.linalg.norm(self.mean_acc)
        if abs(na - 1.0) > 0.1 or np.linalg.norm(self.mean_gyr) > 10.0 / 180.0 * np.pi:
            self.b_first_frame = True
            return
        g = -self.mean_acc / na * G_M_S2
        x.grav = g / np.linalg.norm(g) * E.S2_LEN          # S2(vec): normalised to the manifold's length (S2.hpp:123-126)
        x.vel = self.vel_last.copy()
        x.bg = np.zeros(3); x.ba = np.zeros(3)
        x.offset_T_L_I = self.Lidar_T.copy()
        x.offset_R_L_I = E.R_to_quat(self.Lidar_R)
        P[:] = E.init_P()
        self.last_imu = imu[-1].copy()
        self.last_lidar_end_time = meas["lidar_end_time"]

    def _set_Q(self):
        Q = self.Q
        Q[0:3, 0:3] = np.diag(self.cov_gyr); Q[3:6, 3:6] = np.diag(self.cov_acc)
        Q[6:9, 6:9] = np.diag(self.cov_bias_gyr); Q[9:12, 9:12] = np.diag(self.cov_bias_acc)
        return Q

    def _pose(self, t, x: E.State):
        return dict(t=t, acc=self.acc_s_last.copy(), gyr=self.angvel_last.copy(), vel=x.vel.copy(), pos=x.pos.copy(), rot=E.quat_to_R(x.rot))

    # ---------------------------------------------------------------- UndistortPcl :237-406
    def _undistort(self, meas, x: E.State, P: np.ndarray):
        beg, end = meas["lidar_beg_time"], meas["lidar_end_time"]
        scale = G_M_S2 / np.linalg.norm(self.mean_acc)
        if beg > self.last_lidar_end_time:                 # predict the state at the scan start time
            gyr = self.last_imu[1:4].copy(); acc = self.last_imu[4:7] * scale
            dt = beg - self.last_lidar_end_time
            P[:] = E.predict(x, P, dt, self._set_Q(), acc, gyr)
            self.angvel_last = gyr - x.bg
            self.acc_s_last = E.quat_to_R(x.rot) @ (acc - x.ba) + x.grav
            self.last_lidar_end_time = beg
        v_imu = np.vstack([self.last_imu[None], meas["imu"]])
        imu_end_time = v_imu[-1, 0]
        order = np.argsort(meas["time_ms"], kind="stable")  # std::sort by curvature (unstable in the reference: ties unspecified)
        pts = meas["points"][order].copy(); tms = meas["time_ms"][order].astype(np.float32)
        self.IMUpose = [self._pose(0.0, x)]
        for i in range(len(v_imu) - 1):
            head, tail = v_imu[i], v_imu[i + 1]
            if tail[0] < self.last_lidar_end_time:
                continue
            gyr = 0.5 * (head[1:4] + tail[1:4]); acc = 0.5 * (head[4:7] + tail[4:7]) * scale
            dt = tail[0] - self.last_lidar_end_time if head[0] < self.last_lidar_end_time else tail[0] - head[0]
            dt = min(1.0, dt)
            P[:] = E.predict(x, P, dt, self._set_Q(), acc, gyr)
            self.angvel_last = gyr - x.bg
            self.acc_s_last = E.quat_to_R(x.rot) @ (acc - x.ba) + x.grav
            self.IMUpose.append(self._pose(tail[0] - beg, x))
        gyr = v_imu[-1, 1:4].copy(); acc = v_imu[-1, 4:7] * scale
        note = 1.0 if end > imu_end_time else -1.0
        dt = min(1.0, note * (end - imu_end_time))
        P[:] = E.predict(x, P, dt, self.Q, acc, gyr)
        self.angvel_last = gyr - x.bg
        self.acc_s_last = E.quat_to_R(x.rot) @ (acc - x.ba) + x.grav
        self.IMUpose.append(self._pose(end - beg, x))
        self.last_imu = meas["imu"][-1].copy()
        self.last_lidar_end_time = end
        if not self.undistort or len(pts) == 0:
            return pts
        # backward propagation (:364-405).  double arithmetic, results stored to float32 point fields.
        R_end = E.quat_to_R(x.rot); RL = E.quat_to_R(x.offset_R_L_I); tL = x.offset_T_L_I; p_end = x.pos
        it = len(pts) - 1
        for k in range(len(self.IMUpose) - 1, 0, -1):
            head, tail = self.IMUpose[k - 1], self.IMUpose[k]
            while np.float64(tms[it]) / 1000.0 > head["t"]:
                dt = np.float64(tms[it]) / 1000.0 - head["t"]
                R_i = head["rot"] @ exp_rodrigues(tail["gyr"], dt)
                P_i = pts[it, :3].astype(np.float64)
                T_ei = head["pos"] + head["vel"] * dt + 0.5 * tail["acc"] * dt * dt - p_end
                Pc = RL.T @ (R_end.T @ (R_i @ (RL @ P_i + tL) + T_ei) - tL)
                pts[it, :3] = Pc.astype(np.float32)
                if it == 0:
                    break                                  # NB the first point is re-visited by the next segment (:399)
                it -= 1
        return pts

    # ---------------------------------------------------------------- Process :408-450
    def process(self, meas, x: E.State, P: np.ndarray):
        """Returns the undistorted cloud [n,4] (sorted by time) or None while the IMU is initialising."""
        if len(meas["imu"]) == 0:
            return None
        if self.imu_need_init:
            self._imu_init(meas, x, P)
            self.imu_need_init = True
            self.last_imu = meas["imu"][-1].copy()
            if self.init_iter_num > MAX_INI_COUNT:
                self.imu_need_init = False
                self.cov_acc = self.cov_acc_scale.copy()
                self.cov_gyr = self.cov_gyr_scale.copy()
            return None
        self.state_init_done = True
        return self._undistort(meas, x, P)
