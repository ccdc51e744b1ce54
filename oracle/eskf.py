"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's iterated ESKF update.

Follows, line by line, the parts of IKFoM the LIO hot path executes (paths relative to
/root/reference/slam/mapping/fastlio/include):
  * esekf::update_iterated_dyn_share_modified      IKFoM_toolkit/esekfom/esekfom.hpp:1619-1931
  * state_ikfom (23 DOF / 24 DIM manifold)          use-ikfom.hpp:12-21
  * SO3 boxplus/boxminus/exp/log                    IKFoM_toolkit/mtk/types/SOn.hpp:237-290,
                                                    mtk/src/mtkmath.hpp:117-150,236-275
  * S2 boxplus/boxminus/Bx/Nx_yy/Mx                 IKFoM_toolkit/mtk/types/S2.hpp:132-290
  * A_matrix                                        mtk/src/mtkmath.hpp:222-234

Pin status: PINNED.  IKFoM needs Boost.Preprocessor (absent here); oracle/ref_shim_ikfom re-implements
the dozen macros it uses, and the reference's own esekf / state_ikfom then compile unmodified
(oracle/ref_ikfom.cpp -> oracle/_ref/libref_ikfom.so).  tests/test_oracle_ikfom.py checks this restatement — and the
product's host C++, written independently — against it: manifold operators 1e-13, predict 1e-12, the iterated update
(same evaluation count, state 1e-9, covariance 2e-6 relative).

State layout (DOF index): pos 0:3, rot 3:6, offset_R_L_I 6:9, offset_T_L_I 9:12, vel 12:15,
bg 15:18, ba 18:21, grav 21:23.  Quaternions are (x, y, z, w) like Eigen's coeffs().
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

TOL = 1e-11  # MTK::tolerance<double>()
S2_LEN = 98090.0 / 10000.0  # S2<double, 98090, 10000, 1>
N = 23


def hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def cos_sinc_sqrt(x2: float):
    """mtkmath.hpp:117-150"""
    eps = np.finfo(np.float64).eps
    t2 = np.sqrt(eps)
    tn = np.sqrt(t2)
    if x2 >= tn:
        x = np.sqrt(x2)
        return np.cos(x), np.sin(x) / x
    inv = [1 / 3.0, 1 / 4.0, 1 / 5.0, 1 / 6.0, 1 / 7.0, 1 / 8.0, 1 / 9.0]
    cosi, sinc = 1.0, 1.0
    term = -1 / 2.0 * x2
    for i in range(3):
        cosi += term
        term *= inv[2 * i]
        sinc += term
        term *= -inv[2 * i + 1] * x2
    return cosi, sinc


def mtk_exp(vec, scale):
    """mtkmath.hpp:236-243 -> (w, xyz)"""
    norm2 = float(vec @ vec)
    c, s = cos_sinc_sqrt(scale * scale * norm2)
    return c, s * scale * vec


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_to_R(q):
    """Eigen::QuaternionBase::toRotationMatrix"""
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def R_to_quat(R):
    """Eigen quaternion-from-matrix (Shepperd branch order as in Eigen/src/Geometry/Quaternion.h)"""
    t = np.trace(R)
    q = np.zeros(4)
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def so3_exp(vec, scale=1.0):
    w, v = mtk_exp(np.asarray(vec, np.float64), scale / 2)
    return np.array([v[0], v[1], v[2], w])


def so3_log(q):
    """SOn.hpp:284-288 + mtkmath.hpp:254-275 with scale 2, plus_minus_periodicity true"""
    vec = q[:3]
    w = q[3]
    nv = np.linalg.norm(vec)
    if nv < TOL:
        nv = TOL
    s = 2.0 / nv * np.arctan(nv / w)
    return s * vec


def A_matrix(v):
    """mtkmath.hpp:222-234"""
    sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
    n = np.sqrt(sq)
    if n < TOL:
        return np.eye(3)
    H = hat(v)
    return np.eye(3) + (1 - np.cos(n)) / sq * H + (1 - np.sin(n) / n) / sq * (H @ H)


def s2_Bx(vec):
    """S2.hpp:160-215, S2_typ == 1"""
    L = S2_LEN
    if vec[0] + L > TOL:
        d = L + vec[0]
        res = np.array([[-vec[1], -vec[2]],
                        [L - vec[1] * vec[1] / d, -vec[2] * vec[1] / d],
                        [-vec[2] * vec[1] / d, L - vec[2] * vec[2] / d]])
        return res / L
    res = np.zeros((3, 2))
    res[1, 1] = -1
    res[2, 0] = 1
    return res


def s2_boxplus(vec, delta):
    Bu = s2_Bx(vec) @ delta
    w, v = mtk_exp(Bu, 0.5)
    return quat_to_R(np.array([v[0], v[1], v[2], w])) @ vec


def s2_boxminus(vec, other):
    """S2.hpp:132-157: this (-) other"""
    v_sin = np.linalg.norm(hat(vec) @ other)
    v_cos = float(vec @ other)
    theta = np.arctan2(v_sin, v_cos)
    if v_sin < TOL:
        if abs(theta) > TOL:
            return np.array([3.1415926, 0.0])
        return np.zeros(2)
    Bx = s2_Bx(other)
    return theta / v_sin * (Bx.T @ (hat(other) @ vec))


def s2_Nx_yy(vec):
    return 1 / S2_LEN / S2_LEN * (s2_Bx(vec).T @ hat(vec))


def s2_Mx(vec, delta):
    """S2.hpp:258-272.  NB the reference's `scalar(1/2)` is an integer division = 0, so exp_delta
    is the identity rotation; restated as such."""
    Bx = s2_Bx(vec)
    if np.linalg.norm(delta) < TOL:
        return -hat(vec) @ Bx
    Bu = Bx @ delta
    return -np.eye(3) @ hat(vec) @ A_matrix(Bu).T @ Bx


@dataclass
class State:
    pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rot: np.ndarray = field(default_factory=lambda: np.array([0.0, 0, 0, 1]))  # x,y,z,w
    offset_R_L_I: np.ndarray = field(default_factory=lambda: np.array([0.0, 0, 0, 1]))
    offset_T_L_I: np.ndarray = field(default_factory=lambda: np.zeros(3))
    vel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    bg: np.ndarray = field(default_factory=lambda: np.zeros(3))
    ba: np.ndarray = field(default_factory=lambda: np.zeros(3))
    grav: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -S2_LEN]))

    def copy(self):
        return copy.deepcopy(self)

    def boxplus(self, d):
        self.pos = self.pos + d[0:3]
        self.rot = quat_mul(self.rot, so3_exp(d[3:6]))
        self.offset_R_L_I = quat_mul(self.offset_R_L_I, so3_exp(d[6:9]))
        self.offset_T_L_I = self.offset_T_L_I + d[9:12]
        self.vel = self.vel + d[12:15]
        self.bg = self.bg + d[15:18]
        self.ba = self.ba + d[18:21]
        self.grav = s2_boxplus(self.grav, d[21:23])

    def boxminus(self, o: "State"):
        r = np.zeros(N)
        r[0:3] = self.pos - o.pos
        r[3:6] = so3_log(quat_mul(quat_conj(o.rot), self.rot))
        r[6:9] = so3_log(quat_mul(quat_conj(o.offset_R_L_I), self.offset_R_L_I))
        r[9:12] = self.offset_T_L_I - o.offset_T_L_I
        r[12:15] = self.vel - o.vel
        r[15:18] = self.bg - o.bg
        r[18:21] = self.ba - o.ba
        r[21:23] = s2_boxminus(self.grav, o.grav)
        return r

    def to_vec(self):
        """26 doubles, the layout of include/lsdreg.h lsd_lio_state_t."""
        return np.concatenate([self.pos, self.rot, self.offset_R_L_I, self.offset_T_L_I, self.vel, self.bg,
                               self.ba, self.grav]).astype(np.float64)

    @staticmethod
    def from_vec(v):
        v = np.asarray(v, np.float64)
        return State(v[0:3].copy(), v[3:7].copy(), v[7:11].copy(), v[11:14].copy(), v[14:17].copy(),
                     v[17:20].copy(), v[20:23].copy(), v[23:26].copy())


def init_P():
    """IMU_Processing.hpp:224-230"""
    P = np.eye(N)
    for i in (6, 7, 8, 9, 10, 11):
        P[i, i] = 0.00001
    for i in (15, 16, 17):
        P[i, i] = 0.0001
    for i in (18, 19, 20):
        P[i, i] = 0.001
    P[21, 21] = P[22, 22] = 0.00001
    return P


SO3_IDX = (3, 6)
S2_IDX = 21


def update_iterated(x: State, P: np.ndarray, h_model, R: float = 0.001, maximum_iter: int = 4,
                    limit: float = 0.001, trace: list | None = None, inv=np.linalg.inv):
    """esekfom.hpp:1619-1931.  h_model(state, converge) -> dict(valid, HTH[15,15], HTh[15], n)
    (and 'h_x' [n,15], 'h' [n] when n < 23).  Returns (x, P, n_iters_run)."""
    x = x.copy()
    P = P.copy()
    x_prop = x.copy()
    P_prop = P.copy()
    converge = True
    t = 0
    K_x = np.zeros((N, N))
    iters = 0
    for i in range(-1, maximum_iter):
        m = h_model(x, converge)
        iters += 1
        if not m["valid"]:
            continue
        dof = m["n"]
        dx = x.boxminus(x_prop)
        dx_new = dx.copy()
        P = P_prop.copy()
        for idx in SO3_IDX:
            A = A_matrix(dx[idx:idx + 3]).T
            dx_new[idx:idx + 3] = A @ dx_new[idx:idx + 3]
            P[idx:idx + 3, :] = A @ P[idx:idx + 3, :]
            P[:, idx:idx + 3] = P[:, idx:idx + 3] @ A.T
        seg = dx[S2_IDX:S2_IDX + 2]
        r2 = s2_Nx_yy(x.grav) @ s2_Mx(x_prop.grav, seg)
        dx_new[S2_IDX:S2_IDX + 2] = r2 @ dx_new[S2_IDX:S2_IDX + 2]
        P[S2_IDX:S2_IDX + 2, :] = r2 @ P[S2_IDX:S2_IDX + 2, :]
        P[:, S2_IDX:S2_IDX + 2] = P[:, S2_IDX:S2_IDX + 2] @ r2.T

        if N > dof:
            h_x = np.zeros((dof, N))
            h_x[:, :15] = m["h_x"]
            K = P @ h_x.T @ inv(h_x @ P @ h_x.T / R + np.eye(dof)) / R
            K_h = K @ m["h"]
            K_x = K @ h_x
        else:
            P_temp = inv(P / R)
            P_temp[:15, :15] += m["HTH"]
            P_inv = inv(P_temp)
            K_h = P_inv[:, :15] @ m["HTh"]
            K_x = np.zeros((N, N))
            K_x[:, :15] = P_inv[:, :15] @ m["HTH"]
        dx_ = K_h + (K_x - np.eye(N)) @ dx_new
        x.boxplus(dx_)
        converge = bool(np.all(np.abs(dx_) <= limit))
        if trace is not None:
            trace.append(dict(i=i, dx=dx_.copy(), n=dof, x=x.copy(), converge=converge))
        if converge:
            t += 1
        if (not t) and i == maximum_iter - 2:
            converge = True
        if t > 1 or i == maximum_iter - 1:
            L = P.copy()
            for idx in SO3_IDX:
                A = A_matrix(dx_[idx:idx + 3]).T
                L[idx:idx + 3, :] = A @ P[idx:idx + 3, :]
                K_x[idx:idx + 3, :15] = A @ K_x[idx:idx + 3, :15]
                L[:, idx:idx + 3] = L[:, idx:idx + 3] @ A.T
                P[:, idx:idx + 3] = P[:, idx:idx + 3] @ A.T
            seg = dx_[S2_IDX:S2_IDX + 2]
            r2 = s2_Nx_yy(x.grav) @ s2_Mx(x_prop.grav, seg)
            L[S2_IDX:S2_IDX + 2, :] = r2 @ P[S2_IDX:S2_IDX + 2, :]
            K_x[S2_IDX:S2_IDX + 2, :15] = r2 @ K_x[S2_IDX:S2_IDX + 2, :15]
            L[:, S2_IDX:S2_IDX + 2] = L[:, S2_IDX:S2_IDX + 2] @ r2.T
            P[:, S2_IDX:S2_IDX + 2] = P[:, S2_IDX:S2_IDX + 2] @ r2.T
            P = L - K_x[:, :15] @ P[:15, :]
            return x, P, iters
    return x, P, iters


# ---------------------------------------------------------------------------------------------
# Forward propagation (row N1).  esekf::predict (esekfom.hpp:279-383, dense branch) with the process
# model of use-ikfom.hpp:49-88 (get_f, df_dx, df_dw).  Flattened DIM layout (24): pos 0, rot 3,
# offset_R 6, offset_T 9, vel 12, bg 15, ba 18, grav 21.  The reference evaluates
# MTK::exp(.., scalar_type(1/2)) with an INTEGER 1/2 == 0 (esekfom.hpp:312,344), i.e. the "exp"
# blocks of F_x1 are identities; restated as such.
# ---------------------------------------------------------------------------------------------
def get_f(x: State, acc, gyro):
    f = np.zeros(24)
    f[0:3] = x.vel
    f[3:6] = gyro - x.bg
    f[12:15] = quat_to_R(x.rot) @ (acc - x.ba) + x.grav
    return f


def df_dx(x: State, acc, gyro):
    F = np.zeros((24, 23))
    F[0:3, 12:15] = np.eye(3)
    R = quat_to_R(x.rot)
    F[12:15, 3:6] = -R @ hat(acc - x.ba)
    F[12:15, 18:21] = -R
    F[12:15, 21:23] = s2_Mx(x.grav, np.zeros(2))
    F[3:6, 15:18] = -np.eye(3)
    return F


def df_dw(x: State):
    W = np.zeros((24, 12))
    W[12:15, 3:6] = -quat_to_R(x.rot)
    W[3:6, 0:3] = -np.eye(3)
    W[15:18, 6:9] = np.eye(3)
    W[18:21, 9:12] = np.eye(3)
    return W


def state_oplus(x: State, f, dt):
    """MTK_BUILD_MANIFOLD oplus: vect += f*dt; SO3 *= exp(f, dt) (SOn.hpp:242-245); S2 rotated by exp(f, dt) (S2.hpp:129-134)."""
    x.pos = x.pos + f[0:3] * dt
    x.rot = quat_mul(x.rot, so3_exp(f[3:6], dt))
    x.offset_R_L_I = quat_mul(x.offset_R_L_I, so3_exp(f[6:9], dt))
    x.offset_T_L_I = x.offset_T_L_I + f[9:12] * dt
    x.vel = x.vel + f[12:15] * dt
    x.bg = x.bg + f[15:18] * dt
    x.ba = x.ba + f[18:21] * dt
    x.grav = quat_to_R(so3_exp(f[21:24], dt)) @ x.grav


def predict(x: State, P: np.ndarray, dt: float, Q: np.ndarray, acc, gyro):
    """In place on x; returns the propagated covariance."""
    acc = np.asarray(acc, np.float64); gyro = np.asarray(gyro, np.float64)
    f = get_f(x, acc, gyro)
    fx = df_dx(x, acc, gyro)
    fw = df_dw(x)
    x_before = x.copy()
    state_oplus(x, f, dt)
    F1 = np.eye(N)
    fx_final = np.zeros((N, N)); fw_final = np.zeros((N, 12))
    for idx, dim, dof in ((0, 0, 3), (9, 9, 3), (12, 12, 3), (15, 15, 3), (18, 18, 3)):   # vect states
        fx_final[idx:idx + dof] = fx[dim:dim + dof]
        fw_final[idx:idx + dof] = fw[dim:dim + dof]
    for idx, dim in ((3, 3), (6, 6)):                                                     # SO3 states
        seg = -f[dim:dim + 3] * dt
        A = A_matrix(seg)
        fx_final[idx:idx + 3] = A @ fx[dim:dim + 3]
        fw_final[idx:idx + 3] = A @ fw[dim:dim + 3]
    idx, dim = 21, 21                                                                     # S2 state
    seg = f[dim:dim + 3] * dt
    Nx = s2_Nx_yy(x.grav)
    Mx = s2_Mx(x_before.grav, np.zeros(2))
    F1[idx:idx + 2, idx:idx + 2] = Nx @ Mx
    tmp = -Nx @ hat(x_before.grav) @ A_matrix(seg).T
    fx_final[idx:idx + 2] = tmp @ fx[dim:dim + 3]
    fw_final[idx:idx + 2] = tmp @ fw[dim:dim + 3]
    F1 = F1 + fx_final * dt
    return F1 @ P @ F1.T + (dt * fw_final) @ Q @ (dt * fw_final).T
