"""TEST INFRASTRUCTURE ONLY — CPU restatement of the per-scan LIO driver.

Mirrors fastlio_main (src/laserMapping.cpp:1126-1387) from the point where the undistorted scan
(`feats_undistort`) exists: VoxelGrid 0.5 m -> (first scan: seed the map) -> iterated ESKF update
with h_share_model_geometric -> map_incremental.  IMU propagation/undistortion (row N1 of
SURVEY.md §8f) is outside the hot path: the caller supplies the propagated state/covariance.

Heavy per-point work runs in oracle/lsd_oracle.c (port) and the 23-dim algebra in oracle/eskf.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import eskf
from . import oracle as O


class OracleLio:
    NUM_MAX_ITERATIONS = 4  # laserMapping.cpp:1026
    LASER_POINT_COV = 0.001  # laserMapping.cpp:71
    FILTER_SURF = 0.5  # laserMapping.cpp:1027
    FILTER_MAP = 0.5  # laserMapping.cpp:1028

    def __init__(self, nearby: int = 18, knn_exact: bool = False, expected_cells: int = 1 << 18,
                 nthreads: int = 8, degenerate_detect: bool = True, backend: str = "port", stale_neighbours: bool = False,
                 reference_order: bool = True):
        """backend "port": oracle/lsd_oracle.c; "reference": the compiled reference IVox +
        esti_plane (oracle/_ref) inside the same restated loop.

        stale_neighbours: reproduce the reference's Nearest_Points quirk.  Nearest_Points is a file-scope
        vector<PointVector> that is only resize()d per scan (laserMapping.cpp:1273) and IVox::GetClosestPoint returns
        BEFORE clearing its output when no candidate is in range (ivox3d.h:155-157): a scan point with nothing near it keeps
        the neighbours row i held the last time it found any — from an earlier iteration or an earlier scan — and is
        selected, plane-fitted and gated on those.  False (default) treats such a point as having no neighbours, which is
        what the product does today (known deviation); tests/test_oracle_fastlio.py measures the difference."""
        self.stale = bool(stale_neighbours) and not knn_exact
        self.backend = backend
        if backend == "reference":
            assert O.HAVE_REF and not knn_exact
            self.map = O.RefIvox(0.5, nearby)
            self._hm, self._mi = O.ref.ref_lio_hmodel, O.ref.ref_map_incremental
        else:
            self.map = O.OracleIvox(0.5, nearby, expected_cells)  # laserMapping.cpp:1060-1065
            self._hm, self._mi = O.port.orc_lio_hmodel, O.port.orc_map_incremental
        # reference_order (port backend): the five neighbours in the order IVox::GetClosestPoint leaves them in (libstdc++'s
        # nth_element, ivox3d.h:159-164) instead of ascending (d2, id); esti_plane's fp32 solve depends on the row order.
        self.knn_mode = (1 if knn_exact else 0) | (2 if self.stale else 0) | (4 if reference_order and not knn_exact else 0)
        self._rows = 0  # rows of the persistent neighbour table in use (stale mode): Nearest_Points.size()
        self.nthreads = nthreads
        self.degenerate_detect = degenerate_detect
        self.x = eskf.State()
        self.P = eskf.init_P()
        self.next_id = 0
        self.ekf_inited = True  # flg_EKF_inited, laserMapping.cpp:1196
        self.last = {}

    # -- map seeding (first scan / prebuilt map) : ivox->AddPoints, laserMapping.cpp:1227-1239
    def add_map_points(self, world_xyz: np.ndarray):
        self.map.add(np.ascontiguousarray(world_xyz[:, :4] if self.backend == "reference" else world_xyz[:, :3], np.float32), self.next_id)
        self.next_id += world_xyz.shape[0]

    def _hmodel(self, body, st: eskf.State, converge: bool):
        n = body.shape[0]
        R = np.ascontiguousarray(eskf.quat_to_R(st.rot))
        R_LI = np.ascontiguousarray(eskf.quat_to_R(st.offset_R_L_I))
        HTH6 = np.zeros(36)
        HTh6 = np.zeros(6)
        res_sum = np.zeros(1)
        degen = np.zeros(1, np.int32)
        hx = np.zeros((max(n, 1), 6))
        hh = np.zeros(max(n, 1))
        ne = self._hm(self.map.h, body, n, R, np.ascontiguousarray(st.pos), R_LI,
                                   np.ascontiguousarray(st.offset_T_L_I), int(converge), self.knn_mode,
                                   self.near_xyz, self.near_ids, self.near_cnt, self.selected, self.world,
                                   self.plane, HTH6, HTh6, res_sum, degen, int(self.degenerate_detect),
                                   hx.ctypes.data, hh.ctypes.data, self.nthreads)
        HTH = np.zeros((15, 15))
        HTH[:6, :6] = HTH6.reshape(6, 6)
        HTh = np.zeros(15)
        HTh[:6] = HTh6
        h_x = np.zeros((ne, 15))
        h_x[:, :6] = hx[:ne]
        self.last = dict(n_eff=ne, res_sum=float(res_sum[0]), degenerate=int(degen[0]), HTH6=HTH6.reshape(6, 6).copy(),
                         HTh6=HTh6.copy())
        return dict(valid=ne >= 1, n=ne, HTH=HTH, HTh=HTh, h_x=h_x, h=hh[:ne].copy())

    def process_scan(self, scan: np.ndarray, prior: eskf.State | None = None, P_prior: np.ndarray | None = None,
                     downsample: bool = True, update_map: bool = True, trace: list | None = None):
        """One fastlio_main() pass.  scan: [N,4] undistorted points in the lidar frame."""
        if prior is not None:
            self.x = prior.copy()
        if P_prior is not None:
            self.P = P_prior.copy()
        body = O.voxelgrid(scan, self.FILTER_SURF) if downsample else np.ascontiguousarray(scan, np.float32)
        n = body.shape[0]
        self.body = body
        if self.map.num_cells == 0:  # laserMapping.cpp:1227-1239
            if n > 5:
                w = self._to_world(body, self.x)
                self.add_map_points(w)
            return dict(seeded=True, n_down=n)
        if n < 5:  # laserMapping.cpp:1252-1256
            return dict(seeded=False, n_down=n, skipped=True)
        if not self.stale:
            self.near_xyz = np.zeros((n, 5, 3), np.float32)
            self.near_ids = np.full((n, 5), -1, np.int32)
            self.near_cnt = np.zeros(n, np.int32)
        else:  # Nearest_Points.resize(feats_down_size): rows < n keep their lists, rows >= n are destroyed
            if not hasattr(self, "near_cnt") or self.near_cnt.shape[0] < n:
                cap = max(n, 100000)
                old = getattr(self, "near_cnt", None)
                nx, ni, nc = np.zeros((cap, 5, 3), np.float32), np.full((cap, 5), -1, np.int32), np.zeros(cap, np.int32)
                if old is not None:
                    k = old.shape[0]
                    nx[:k], ni[:k], nc[:k] = self.near_xyz, self.near_ids, self.near_cnt
                self.near_xyz, self.near_ids, self.near_cnt = nx, ni, nc
            if self._rows > n:
                self.near_cnt[n:self._rows] = 0
                self.near_ids[n:self._rows] = -1
            self._rows = n
        if not hasattr(self, "selected") or self.selected.shape[0] < n:
            self.selected = np.ones(max(n, 100000), np.uint8)  # memset(true), laserMapping.cpp:1089
        self.world = np.zeros((n, 4), np.float32)
        self.plane = np.zeros((n, 4), np.float32)
        iters_log = []

        def hm(st, conv):
            r = self._hmodel(body, st, conv)
            iters_log.append(dict(self.last, converge_in=conv))
            return r

        self.x, self.P, iters = eskf.update_iterated(self.x, self.P, hm, self.LASER_POINT_COV,
                                                     self.NUM_MAX_ITERATIONS, 0.001, trace)
        added = 0
        if update_map:
            added = self.map_incremental(body)
        return dict(seeded=False, n_down=n, iters=iters, log=iters_log, added=added)

    def _to_world(self, body, st):
        R = eskf.quat_to_R(st.rot)
        R_LI = eskf.quat_to_R(st.offset_R_L_I)
        p = body[:, :3].astype(np.float64)
        w = (R @ (R_LI @ p.T + st.offset_T_L_I[:, None]) + st.pos[:, None]).T
        out = body.copy()
        out[:, :3] = w.astype(np.float32)
        return out

    def map_incremental(self, body):
        n = body.shape[0]
        st = self.x
        R = np.ascontiguousarray(eskf.quat_to_R(st.rot))
        R_LI = np.ascontiguousarray(eskf.quat_to_R(st.offset_R_L_I))
        self.flags = np.zeros(n, np.uint8)
        # reference_order: ids grow in the reference's insertion order (all PointToAdd of the scan, then all PointNoNeedDownsample,
        # laserMapping.cpp:571-572) — the product orders a voxel's points by id to rebuild the reference's candidate sequence
        id_t2 = n if (self.knn_mode & 4) else 0
        added = self._mi(self.map.h, body, n, R, np.ascontiguousarray(st.pos), R_LI,
                                           np.ascontiguousarray(st.offset_T_L_I), self.near_xyz, self.near_cnt,
                                           int(self.ekf_inited), self.FILTER_MAP, self.world, self.flags, self.next_id, id_t2)
        self.next_id += n + id_t2
        return added
