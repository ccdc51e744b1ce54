"""TEST INFRASTRUCTURE ONLY — the whole LIO front-end, twice.

RefFastLio     the UNMODIFIED reference pipeline (slam/mapping/fastlio/src/laserMapping.cpp: fastlio_init / _imu_enqueue /
               _pcl_enqueue / fastlio_main, with preprocess.cpp, ImuProcess, the IKFoM filter, iVox, esti_plane),
               compiled by oracle/Makefile into oracle/_ref/libref_fastlio.so and driven through oracle/ref_fastlio.cpp.
               Of its stages only pcl::VoxelGrid is a restatement (PCL is external to the reference tree).
OracleFastLio  the same pipeline assembled from this directory's restatements (imu.py, eskf.py, lio.py, lsd_oracle.c),
               following fastlio_main line by line (laserMapping.cpp:1126-1387).

tests/test_oracle_fastlio.py runs both on the same sensor stream; agreement scan by scan pins every restated stage AND
their composition (state hand-over between ImuProcess and the filter, flg_EKF_inited, the NEARBY74 -> NEARBY18 switch,
the seeding scan) against the reference itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import eskf as E
from .imu import OracleImuProcess
from .lio import OracleLio

_HERE = os.path.dirname(os.path.abspath(__file__))
# LSD_REF_FASTLIO_LIB: another build of the same unmodified sources (tools/ref_build_sensitivity.py compiles the reference
# with other compiler flags to measure how far the reference's own poses move between its builds)
_PATH = os.environ.get("LSD_REF_FASTLIO_LIB") or os.path.join(_HERE, "_ref", "libref_fastlio.so")
HAVE_REF_FASTLIO = os.path.exists(_PATH)
_lib = None

INIT_TIME = 0.1      # laserMapping.cpp:70
BLIND = 0.1          # laserMapping.cpp:1094


def _load():
    global _lib
    if _lib is None:
        if not HAVE_REF_FASTLIO:
            raise RuntimeError("oracle/_ref/libref_fastlio.so missing (built only where /root/reference exists)")
        d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        u = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
        L = C.CDLL(_PATH)
        L.ref_fastlio_init.restype = C.c_int
        L.ref_fastlio_init.argtypes = [d, d, C.c_int, C.c_int, C.c_double, C.c_int]
        L.ref_fastlio_imu.argtypes = [C.c_double, d, d]
        L.ref_fastlio_scan.argtypes = [f, u, C.c_int, C.c_uint64]
        L.ref_fastlio_main.restype = C.c_int
        L.ref_fastlio_is_init.restype = C.c_int
        L.ref_fastlio_get_state.argtypes = [d, C.c_void_p]
        L.ref_fastlio_counts.argtypes = [C.c_void_p] * 4
        L.ref_fastlio_get_down.restype = C.c_int
        L.ref_fastlio_get_down.argtypes = [C.c_void_p, C.c_int]
        if hasattr(L, "ref_fastlio_pass"):
            L.ref_fastlio_set_threads.argtypes = [C.c_int]
            L.ref_fastlio_bench_init.restype = C.c_int
            L.ref_fastlio_bench_init.argtypes = [C.c_size_t]
            L.ref_fastlio_map_add.argtypes = [f, C.c_int, C.c_int]
            L.ref_fastlio_pass.restype = C.c_int
            L.ref_fastlio_pass.argtypes = [f, C.c_int, d, d]
        _lib = L
    return _lib


class RefFastLio:
    """One instance per process: the reference keeps the pipeline in file-scope globals (fastlio_init resets them)."""

    def __init__(self, ext_R=None, ext_t=None, filter_num=1, max_point_num=-1, scan_period=0.1, undistort=True):
        self.L = _load()
        R = np.ascontiguousarray(np.eye(3) if ext_R is None else ext_R, np.float64).reshape(-1)
        t = np.ascontiguousarray(np.zeros(3) if ext_t is None else ext_t, np.float64)
        self.L.ref_fastlio_init(t, R, filter_num, max_point_num, scan_period, int(undistort))

    def push_imu(self, stamp, gyr, acc_g):
        """acc in units of g, as the rest of this directory; the reference entry takes m/s^2 and divides by 9.81."""
        self.L.ref_fastlio_imu(float(stamp), np.ascontiguousarray(gyr, np.float64), np.ascontiguousarray(acc_g, np.float64) * 9.81)

    def push_scan(self, xyzi, stamp_us, header_stamp_us):
        xyzi = np.ascontiguousarray(xyzi, np.float32)
        self.L.ref_fastlio_scan(xyzi, np.ascontiguousarray(stamp_us, np.uint32), xyzi.shape[0], int(header_stamp_us))

    def step(self):
        return bool(self.L.ref_fastlio_main())

    @property
    def initialised(self):
        return bool(self.L.ref_fastlio_is_init())

    def state(self):
        x = np.zeros(26); P = np.zeros((23, 23))
        self.L.ref_fastlio_get_state(x, P.ctypes.data)
        return E.State.from_vec(x), P

    def counts(self):
        v = (C.c_int * 4)()
        self.L.ref_fastlio_counts(*[C.byref(v, 4 * i) for i in range(4)])
        return dict(n_down=v[0], n_eff=v[1], map_cells=v[2], degenerate=v[3])

    def downsampled(self):
        n = self.L.ref_fastlio_get_down(None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.ref_fastlio_get_down(out.ctypes.data, n)
        return out[:n]


class RefFastLioBench:
    """bench.py's CPU arm: the reference's per-scan hot path (VoxelGrid -> kf.update_iterated_dyn_share_modified with
    h_share_model -> map_incremental: the statements of fastlio_main after ImuProcess, executed by the compiled reference
    objects, oracle/ref_fastlio.cpp::ref_fastlio_pass) on a prebuilt map.  One instance per process (file-scope state)."""

    def __init__(self, capacity=1 << 30, threads=8):
        self.L = _load()
        if not hasattr(self.L, "ref_fastlio_pass"):
            raise RuntimeError("oracle/_ref/libref_fastlio.so predates ref_fastlio_pass: rebuild it (make -C oracle ref)")
        self.L.ref_fastlio_bench_init(capacity)
        self.set_threads(threads)

    def set_threads(self, n):
        self.threads = int(n)
        self.L.ref_fastlio_set_threads(int(n))

    def add_map_points(self, pts, chunk=1 << 20):
        pts = np.ascontiguousarray(pts, np.float32)
        for lo in range(0, pts.shape[0], chunk):
            blk = np.ascontiguousarray(pts[lo:lo + chunk])
            self.L.ref_fastlio_map_add(blk, blk.shape[0], blk.shape[1])

    def process_scan(self, scan, prior: E.State, P: np.ndarray):
        """-> (posterior State, posterior P, feats_down_size or -1 if the reference skipped the scan)"""
        scan = np.ascontiguousarray(scan[:, :4], np.float32)
        x = prior.to_vec(); P = np.ascontiguousarray(P, np.float64).copy()
        n = self.L.ref_fastlio_pass(scan, scan.shape[0], x, P)
        return E.State.from_vec(x), P, n

    def counts(self):
        v = (C.c_int * 4)()
        self.L.ref_fastlio_counts(*[C.byref(v, 4 * i) for i in range(4)])
        return dict(n_down=v[0], n_eff=v[1], map_cells=v[2], degenerate=v[3])


class OracleFastLio:
    """fastlio_main restated on top of OracleImuProcess + OracleLio.  Same feeding protocol as RefFastLio."""

    def __init__(self, ext_R=None, ext_t=None, filter_num=1, max_point_num=-1, scan_period=0.1, undistort=True, nthreads=8,
                 backend="port", stale_neighbours=True, reference_order=True):
        self.imu_proc = OracleImuProcess(ext_R, ext_t, undistort=undistort)          # laserMapping.cpp:1101-1106
        self.lio = OracleLio(nearby=74, nthreads=nthreads, backend=backend,          # NEARBY74 first, laserMapping.cpp:1062
                             stale_neighbours=stale_neighbours, reference_order=reference_order)
        self.lio.x, self.lio.P = E.State(), np.eye(23)                               # a default-constructed esekf:
        self.lio.x.grav = np.array([E.S2_LEN, 0.0, 0.0])                             # S2() = length * e_x (S2.hpp:62-66) until IMU_init
        self.filter_num, self.max_point_num, self.scan_period = filter_num, max_point_num, scan_period
        self.imu_buf, self.scan_buf = [], []
        self.first_scan, self.first_lidar_time = True, 0.0
        self.nearby = 74
        self.last = {}

    def push_imu(self, stamp, gyr, acc_g):
        # acc * 9.81 / 9.81 as the reference's driver + fastlio_imu_enqueue do (laserMapping.cpp:414): not always a no-op in fp64
        a = np.asarray(acc_g, np.float64) * 9.81 / 9.81
        self.imu_buf.append(np.concatenate([[float(stamp)], np.asarray(gyr, np.float64), a]))

    def push_scan(self, xyzi, stamp_us, header_stamp_us):
        """Preprocess::velodyne_handler (preprocess.cpp:280-427): decimate, drop the blind zone, time in ms as fp32."""
        xyzi = np.ascontiguousarray(xyzi, np.float32)
        n = xyzi.shape[0]
        step = self.filter_num
        if self.max_point_num > 0:
            step = max(1, n // self.max_point_num)
        keep = (np.arange(n) % step) == 0
        r2 = xyzi[:, 0] * xyzi[:, 0] + xyzi[:, 1] * xyzi[:, 1] + xyzi[:, 2] * xyzi[:, 2]
        keep &= r2.astype(np.float64) > BLIND * BLIND
        t_ms = (np.asarray(stamp_us, np.uint32).astype(np.float32) / np.float32(1000.0))
        self.scan_buf.append((xyzi[keep], t_ms[keep], int(header_stamp_us) / 1000000.0))

    def _sync(self):  # sync_packages, laserMapping.cpp:445-520
        if not self.scan_buf or not self.imu_buf:
            return None
        pts, t_ms, beg = self.scan_buf.pop(0)
        end = beg + self.scan_period
        k = 0
        while k < len(self.imu_buf) and not (self.imu_buf[k][0] > end):
            k += 1
        imu = np.array(self.imu_buf[:k]).reshape(-1, 7)
        del self.imu_buf[:k]
        return dict(lidar_beg_time=beg, lidar_end_time=end, points=pts, time_ms=t_ms, imu=imu, ins_vel=None)

    def step(self, teacher=None, down_from=None):
        """One fastlio_main() pass.  teacher: optional callable -> (State, P), the reference's posterior of this scan; when
        given, this pipeline's own posterior is kept in self.free_posterior and replaced by the teacher's before
        map_incremental, so that a comparison measures ONE scan's deviation instead of the chaotic growth of fp32
        rounding differences through the map (a single map point that lands in another voxel moves later poses by 1e-5 m).
        down_from: optional callable(undistorted [n,4], prior State, prior P, nearby, ekf_inited) -> downsampled [m,4]; when
        given it replaces this pipeline's VoxelGrid (a GPU parity test runs the product's whole scan there and hands back the
        product's downsampled cloud, so both sides search with identical query points)."""
        self.last = {}
        self.free_posterior = None
        meas = self._sync()
        if meas is None:
            return False
        if self.first_scan:                                                          # :1171-1177
            self.first_lidar_time = meas["lidar_beg_time"]
            self.imu_proc.first_lidar_time = self.first_lidar_time
            self.first_scan = False
            return True
        und = self.imu_proc.process(meas, self.lio.x, self.lio.P)                    # :1188
        if und is None or und.shape[0] == 0:                                         # :1192-1196
            return True
        since = meas["lidar_beg_time"] - self.first_lidar_time
        self.lio.ekf_inited = not (since < INIT_TIME)                                # :1198
        if self.lio.map.num_cells > 0 and self.nearby != 18 and since > 10 * INIT_TIME:   # :1241-1243 (after the seeding return)
            self.lio.map.set_nearby(18)
            self.nearby = 18
        if down_from is not None:
            body = down_from(und, self.lio.x.copy(), self.lio.P.copy(), self.nearby, self.lio.ekf_inited)
            r = self.lio.process_scan(body, downsample=False, update_map=teacher is None)
        else:
            r = self.lio.process_scan(und, downsample=True, update_map=teacher is None)
        self.last = r
        if teacher is not None and "iters" in r:
            self.free_posterior = (self.lio.x.copy(), self.lio.P.copy())
            x, P = teacher()
            self.lio.x, self.lio.P = x.copy(), P.copy()
            r["added"] = self.lio.map_incremental(self.lio.body)
        return True

    @property
    def initialised(self):
        return bool(self.imu_proc.state_init_done)      # ImuProcess::IsInit, IMU_Processing.hpp:125-128

    def state(self):
        return self.lio.x, self.lio.P

    def counts(self):
        n_eff = self.last["log"][-1]["n_eff"] if self.last.get("log") else 0
        degen = self.last["log"][-1]["degenerate"] if self.last.get("log") else 0
        return dict(n_down=self.last.get("n_down", 0), n_eff=n_eff, map_cells=self.lio.map.num_cells, degenerate=degen)

    def downsampled(self):
        return self.lio.body
