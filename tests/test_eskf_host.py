"""CPU: the product's host-side ESKF (csrc/eskf.hpp, through the C ABI lsd_eskf_update_table /
lsd_state_box*) against the independent numpy restatement of IKFoM (oracle/eskf.py)."""
import numpy as np
import pytest

import lsdreg
from oracle import eskf


def _rand_state(rng):
    s = eskf.State()
    s.pos = rng.uniform(-50, 50, 3)
    s.rot = eskf.so3_exp(rng.uniform(-1, 1, 3))
    s.offset_R_L_I = eskf.so3_exp(rng.uniform(-0.1, 0.1, 3))
    s.offset_T_L_I = rng.uniform(-0.5, 0.5, 3)
    s.vel = rng.uniform(-5, 5, 3)
    s.bg = rng.uniform(-0.01, 0.01, 3)
    s.ba = rng.uniform(-0.1, 0.1, 3)
    g = rng.normal(size=3); s.grav = g / np.linalg.norm(g) * eskf.S2_LEN
    return s


def test_boxplus_boxminus_match_oracle_and_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = _rand_state(rng)
        d = rng.uniform(-0.2, 0.2, 23)
        b = a.copy(); b.boxplus(d)
        gb = lsdreg.state_boxplus(a.to_vec(), d)
        np.testing.assert_allclose(gb, b.to_vec(), rtol=0, atol=1e-13)
        r_o = b.boxminus(a)
        r_g = lsdreg.state_boxminus(b.to_vec(), a.to_vec())
        np.testing.assert_allclose(r_g, r_o, rtol=0, atol=1e-12)
        np.testing.assert_allclose(r_g[:21], d[:21], atol=1e-9)   # (a [+] d) [-] a = d on vect / SO3 parts


def _table(rng, n_tab=5):
    HTH, HTh = [], []
    for _ in range(n_tab):
        J = rng.normal(size=(4000, 6)) * np.array([1, 1, 1, 20, 20, 20])
        r = rng.normal(size=4000) * 0.02 + 0.03
        HTH.append(J.T @ J); HTh.append(J.T @ r)
    return np.array(HTH), np.array(HTh)


def _oracle_run(x0, P0, HTH, HTh, n_eff, inv=np.linalg.inv):
    e = [0]

    def hm(st, conv):
        k = min(e[0], len(n_eff) - 1); e[0] += 1
        H15 = np.zeros((15, 15)); H15[:6, :6] = HTH[k]
        h15 = np.zeros(15); h15[:6] = HTh[k]
        return dict(valid=n_eff[k] >= 1, n=max(int(n_eff[k]), 23), HTH=H15, HTh=h15)

    return eskf.update_iterated(x0, P0, hm, inv=inv)


def test_literal_path_matches_numpy_oracle():
    rng = np.random.default_rng(1)
    for trial in range(10):
        x0 = _rand_state(rng)
        P0 = eskf.init_P() + np.diag(rng.uniform(0, 1e-3, 23))
        HTH, HTh = _table(rng)
        n_eff = np.array([4000, 4000, 0, 4000, 4000], np.int32) if trial == 3 else np.full(5, 4000, np.int32)
        xo, Po, it_o = _oracle_run(x0, P0, HTH, HTh, n_eff)
        xg, Pg, it_g = lsdreg.eskf_update_table(x0.to_vec(), P0, HTH, HTh, n_eff, literal=True)
        assert it_g == it_o
        np.testing.assert_allclose(xg, xo.to_vec(), rtol=0, atol=1e-9)
        np.testing.assert_allclose(Pg, Po, rtol=2e-3, atol=1e-13)   # both literal forms lose ~cond*eps in P


def _inv_longdouble(A):
    """Gauss-Jordan with partial pivoting in 80-bit long double: the 'exact' yardstick."""
    n = A.shape[0]
    M = np.concatenate([A.astype(np.longdouble), np.eye(n, dtype=np.longdouble)], 1)
    for k in range(n):
        p = k + int(np.argmax(np.abs(M[k:, k])))
        M[[k, p]] = M[[p, k]]
        M[k] /= M[k, k]
        for i in range(n):
            if i != k:
                M[i] -= M[i, k] * M[k]
    return M[:, n:]


def test_schur_path_matches_literal_and_extended_precision():
    """Default product path (two 6x6 inverses instead of two 23x23): state equal to the literal
    evaluation to 1e-9, covariance equal to the extended-precision evaluation of the same filter
    to 1e-6 relative."""
    rng = np.random.default_rng(2)
    for trial in range(10):
        x0 = _rand_state(rng)
        P0 = eskf.init_P()
        HTH, HTh = _table(rng)
        n_eff = np.full(5, 4000, np.int32)
        xo, Po, _ = _oracle_run(x0, P0, HTH, HTh, n_eff)                       # literal, float64
        xe, Pe, _ = _oracle_run(x0, P0, HTH, HTh, n_eff, inv=lambda A: _inv_longdouble(A).astype(np.float64))
        xg, Pg, _ = lsdreg.eskf_update_table(x0.to_vec(), P0, HTH, HTh, n_eff, literal=False)
        np.testing.assert_allclose(xg, xo.to_vec(), rtol=0, atol=1e-9)
        err_fast = np.abs(np.diag(Pg) - np.diag(Pe)) / np.abs(np.diag(Pe))
        assert err_fast.max() < 1e-6
        np.testing.assert_allclose(Pg, Pe, rtol=1e-4, atol=1e-13)
        assert (np.linalg.eigvalsh(0.5 * (Pg + Pg.T)) > -1e-15).all()


def test_invalid_measurements_leave_state_untouched():
    x0 = eskf.State().to_vec(); P0 = eskf.init_P()
    HTH = np.zeros((1, 36)); HTh = np.zeros((1, 6))
    x, P, ev = lsdreg.eskf_update_table(x0, P0, HTH, HTh, np.zeros(1, np.int32))
    assert ev == 5                                   # i = -1 .. maximum_iter-1, all `continue`d (esekfom.hpp:1633-1641)
    np.testing.assert_array_equal(x, x0); np.testing.assert_array_equal(P, P0)


def test_predict_matches_oracle():
    """esekf::predict (host C++, csrc/imu.cu) against the numpy restatement (oracle/eskf.py::predict):
    state and covariance after chains of IMU steps from random states."""
    import lsdreg
    from oracle import eskf as E
    rng = np.random.default_rng(11)
    Q = np.diag([0.1] * 3 + [0.1] * 3 + [1e-4] * 3 + [1e-4] * 3)
    for trial in range(5):
        x = E.State()
        x.boxplus(rng.normal(0, 0.3, 23))
        x.vel = rng.normal(0, 2.0, 3)
        A = rng.normal(0, 0.05, (23, 23))
        P = E.init_P() + A @ A.T
        xv = x.to_vec()
        Pv = P.copy()
        for step in range(12):
            acc = np.array([0.1, -0.2, 9.7]) + rng.normal(0, 0.5, 3)
            gyr = rng.normal(0, 0.4, 3)
            dt = float(rng.uniform(0.001, 0.02))
            P = E.predict(x, P, dt, Q, acc, gyr)
            xv, Pv = lsdreg.eskf_predict(xv, Pv, dt, Q, acc, gyr)
        np.testing.assert_allclose(xv, x.to_vec(), rtol=0, atol=1e-12)
        np.testing.assert_allclose(Pv, P, rtol=1e-11, atol=1e-14)
        assert np.linalg.eigvalsh(0.5 * (Pv + Pv.T)).min() > 0
