"""Pins the filter and the IMU stage against the COMPILED reference: IKFoM's esekf (predict,
update_iterated_dyn_share_modified), state_ikfom's manifold operators and ImuProcess (IMU_init / UndistortPcl), built from
the reference's own headers with a from-scratch Boost.Preprocessor shim (oracle/ref_ikfom.cpp -> oracle/_ref/libref_ikfom.so).
Checked here: the numpy restatement (oracle/eskf.py, oracle/imu.py) AND the product's host C++ (lsd_eskf_predict,
lsd_eskf_update_table, lsd_state_boxplus/boxminus — no GPU needed)."""
import numpy as np
import pytest

import lsdreg
from oracle import eskf as E
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.HAVE_REF_IKFOM, reason="oracle/_ref/libref_ikfom.so not built (needs /root/reference)")
R = O.ref_ikfom


def _random_state(rng, spread=0.3):
    x = E.State()
    x.boxplus(rng.normal(0, spread, 23))
    x.vel = rng.normal(0, 2.0, 3)
    return x


def test_manifold_operators_match_reference():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = _random_state(rng), _random_state(rng)
        d = rng.normal(0, 0.2, 23)
        xr = a.to_vec().copy(); R.ref_ikfom_boxplus(xr, np.ascontiguousarray(d))
        xo = a.copy(); xo.boxplus(d)
        np.testing.assert_allclose(xo.to_vec(), xr, rtol=0, atol=1e-13)
        np.testing.assert_allclose(lsdreg.state_boxplus(a.to_vec(), d), xr, rtol=0, atol=1e-13)
        dr = np.zeros(23); R.ref_ikfom_boxminus(a.to_vec(), b.to_vec(), dr)
        np.testing.assert_allclose(a.boxminus(b), dr, rtol=0, atol=1e-12)
        np.testing.assert_allclose(lsdreg.state_boxminus(a.to_vec(), b.to_vec()), dr, rtol=0, atol=1e-12)


def test_predict_matches_reference():
    rng = np.random.default_rng(1)
    Q = np.diag([0.1] * 3 + [0.1] * 3 + [1e-4] * 3 + [1e-4] * 3)
    for trial in range(5):
        x = _random_state(rng)
        A = rng.normal(0, 0.05, (23, 23)); P = E.init_P() + A @ A.T
        xr, Pr = x.to_vec().copy(), P.copy()
        xg, Pg = x.to_vec().copy(), P.copy()
        for step in range(15):
            acc = np.array([0.1, -0.2, 9.7]) + rng.normal(0, 0.5, 3); gyr = rng.normal(0, 0.4, 3); dt = float(rng.uniform(0.001, 0.02))
            P = E.predict(x, P, dt, Q, acc, gyr)
            R.ref_ikfom_predict(xr, Pr, dt, np.ascontiguousarray(Q), np.ascontiguousarray(acc), np.ascontiguousarray(gyr))
            xg, Pg = lsdreg.eskf_predict(xg, Pg, dt, Q, acc, gyr)
        for xx, PP in ((x.to_vec(), P), (xg, Pg)):
            np.testing.assert_allclose(xx, xr, rtol=0, atol=1e-12)
            np.testing.assert_allclose(PP, Pr, rtol=1e-10, atol=1e-14)


def _measurements(rng, n_table, n_rows, scale_h):
    rows = np.zeros((n_table, n_rows, 6)); h = np.zeros((n_table, n_rows))
    for e in range(n_table):
        nrm = rng.normal(0, 1, (n_rows, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        p = rng.uniform(-30, 30, (n_rows, 3))
        rows[e, :, :3] = nrm; rows[e, :, 3:] = np.cross(p, nrm)
        h[e] = rng.normal(0, scale_h, n_rows)
    return rows, h


@pytest.mark.parametrize("scale_h,invalid_first,n_rows", [(0.05, False, 400), (1e-4, False, 400), (0.05, True, 400), (0.02, False, 12)])
def test_iterated_update_matches_reference(scale_h, invalid_first, n_rows):
    """update_iterated_dyn_share_modified driven by a tabulated measurement model: same number of evaluations, same state
    and covariance — for the numpy restatement and for the product's Schur-complement and literal two-inversion forms.
    n_rows = 12 < 23 exercises the small-measurement branch (esekfom.hpp:1727)."""
    rng = np.random.default_rng(7)
    n_table = 6
    for trial in range(4):
        x0 = _random_state(rng, 0.1)
        A = rng.normal(0, 0.02, (23, 23)); P0 = E.init_P() + A @ A.T
        rows, h = _measurements(rng, n_table, n_rows, scale_h)
        n_eff = np.full(n_table, n_rows, np.int32)
        if invalid_first:
            n_eff[0] = 0                                              # "No Effective Points" on the first evaluation: skipped
        xr, Pr = x0.to_vec().copy(), P0.copy()
        conv_r = np.full(16, -1, np.int32)
        ev_r = R.ref_ikfom_update_rows(xr, Pr, np.ascontiguousarray(rows), np.ascontiguousarray(h), n_eff, n_table, n_rows, 0.001, 4, 0.001,
                                       conv_r.ctypes.data)
        HTH = np.einsum("eni,enj->eij", rows, rows); HTh = np.einsum("eni,en->ei", rows, h)
        if n_rows >= 23:                                              # the table form of the product carries no h_x rows
            for literal in (False, True):
                conv_g = np.full(16, -1, np.int32)
                xg, Pg, ev_g = lsdreg.eskf_update_table(x0.to_vec(), P0, HTH, HTh, n_eff, R=0.001, max_iterations=4, eps=0.001, literal=literal,
                                                        converge_log=conv_g)
                assert ev_g == ev_r and list(conv_g[:ev_g]) == list(conv_r[:ev_r])   # same search / reuse pattern as the reference filter
                np.testing.assert_allclose(xg, xr, rtol=0, atol=1e-9)
                np.testing.assert_allclose(Pg, Pr, rtol=2e-6, atol=2e-6 * np.abs(Pr).max())   # two ill-conditioned 23x23 inversions in the reference
        calls = [0]
        conv_o = []

        def hm(state, converge):
            conv_o.append(int(converge))                              # the flag that decides on a new neighbour search
            e = min(calls[0], n_table - 1); calls[0] += 1
            if n_eff[e] < 1:
                return dict(valid=False)
            H15 = np.zeros((15, 15)); H15[:6, :6] = HTH[e]
            h15 = np.zeros(15); h15[:6] = HTh[e]
            Hx = np.zeros((n_rows, 15)); Hx[:, :6] = rows[e]
            return dict(valid=True, n=n_rows, HTH=H15, HTh=h15, h_x=Hx, h=h[e])
        xo, Po, _ = E.update_iterated(x0, P0, hm, R=0.001, maximum_iter=4, limit=0.001)
        assert calls[0] == ev_r
        assert conv_o == list(conv_r[:ev_r])
        np.testing.assert_allclose(xo.to_vec(), xr, rtol=0, atol=1e-9)
        np.testing.assert_allclose(Po, Pr, rtol=2e-6, atol=2e-6 * np.abs(Pr).max())


class _RefImu:
    def __init__(self, ext_R, ext_t, undistort=True):
        self.h = R.ref_imu_create(np.ascontiguousarray(ext_R, np.float64).reshape(9), np.ascontiguousarray(ext_t, np.float64), 0.1, 0.1, 1e-4, 1e-4, int(undistort))

    def __del__(self):
        if getattr(self, "h", None):
            R.ref_imu_destroy(self.h); self.h = None

    def process(self, m, x26, P):
        pts = np.ascontiguousarray(m["points"], np.float32); tms = np.ascontiguousarray(m["time_ms"], np.float32)
        out = np.zeros_like(pts); out_t = np.zeros_like(tms)
        imu = np.ascontiguousarray(m["imu"], np.float64)
        n = R.ref_imu_process(self.h, imu, imu.shape[0], None, float(m["lidar_beg_time"]), float(m["lidar_end_time"]), pts, tms, pts.shape[0], x26, P, out, out_t)
        return out[:n], out_t[:n]

    def poses(self):
        buf = np.zeros((256, 22))
        n = R.ref_imu_get_poses(self.h, buf, 256)
        return buf[:n]


@pytest.mark.parametrize("ext,t_min_ms", [(False, 0.0), (True, 0.7)])
def test_imu_process_matches_compiled_reference(ext, t_min_ms):
    """ImuProcess::Process over a whole stream (initialisation, then moving): state, covariance, IMU pose list and every
    undistorted point of oracle/imu.py against the compiled reference class.  t_min_ms > 0 exercises the reference's
    repeated compensation of the earliest point (IMU_Processing.hpp:399)."""
    from lsdreg import synth
    from oracle.imu import OracleImuProcess
    rng = np.random.default_rng(5)
    ext_R = synth.rot_from_rpy(0.02, -0.01, 0.05) if ext else np.eye(3)
    ext_t = np.array([0.3, -0.1, 0.2]) if ext else np.zeros(3)
    o = OracleImuProcess(ext_R=ext_R, ext_t=ext_t)
    r = _RefImu(ext_R, ext_t)
    xo, Po = E.State(), E.init_P()
    xr, Pr = xo.to_vec().copy(), Po.copy()
    per, n_pts, checked = 10, 3000, 0
    for f in range(15):
        beg = 0.1 * f
        moving = f >= 11
        imu = np.zeros((per, 7)); imu[:, 0] = beg + (np.arange(per) + 1) * (0.1 / per)
        imu[:, 1:4] = rng.normal(0, 0.002, (per, 3)) + (np.array([0.0, 0.0, 0.3]) if moving else 0.0)
        imu[:, 4:7] = np.array([0.0, 0.0, 1.0]) + rng.normal(0, 0.002, (per, 3)) + (np.array([0.05, 0.0, 0.0]) if moving else 0.0)
        pts = np.zeros((n_pts, 4), np.float32); pts[:, :3] = rng.uniform(-40, 40, (n_pts, 3)); pts[:, 3] = rng.uniform(0, 255, n_pts)
        tms = np.sort(rng.choice(np.arange(int(t_min_ms * 100) + 1, 10000), n_pts, replace=False)).astype(np.float32) / 100.0   # unique times
        if t_min_ms == 0.0:
            tms[0] = 0.0
        tms = tms[rng.permutation(n_pts)]
        m = dict(lidar_beg_time=beg, lidar_end_time=beg + 0.1, points=pts, time_ms=tms, imu=imu, ins_vel=None)
        out_o = o.process(m, xo, Po)
        out_r, t_r = r.process(m, xr, Pr)
        np.testing.assert_allclose(xo.to_vec(), xr, rtol=0, atol=1e-11)
        np.testing.assert_allclose(Po, Pr, rtol=1e-9, atol=1e-14)
        if out_o is None:
            assert len(out_r) == 0
            continue
        checked += 1
        pr = r.poses()
        assert len(pr) == len(o.IMUpose)
        for k, (a, b) in enumerate(zip(pr, o.IMUpose)):
            np.testing.assert_allclose(a[0], b["t"], atol=1e-12)
            if k > 0 or f > 11:   # the reference never initialises acc_s_last: IMUpose[0].acc of the first scan is garbage (and unused, :379)
                np.testing.assert_allclose(a[1:4], b["acc"], atol=1e-10)
            np.testing.assert_allclose(a[4:7], b["gyr"], atol=1e-12); np.testing.assert_allclose(a[7:10], b["vel"], atol=1e-11)
            np.testing.assert_allclose(a[10:13], b["pos"], atol=1e-11); np.testing.assert_allclose(a[13:22].reshape(3, 3), b["rot"], atol=1e-12)
        assert out_r.shape == out_o.shape
        np.testing.assert_array_equal(t_r, np.sort(tms))                       # both sorted by (unique) time
        d = np.abs(out_r[:, :3].astype(np.float64) - out_o[:, :3])
        assert d.max() <= 8e-6 and (d == 0).mean() > 0.99, (d.max(), (d == 0).mean())
        np.testing.assert_array_equal(out_r[:, 3], out_o[:, 3])
    assert checked >= 4
