"""Seeded synthetic scenes the block generator (lsdreg.synth) does not cover.  Test infrastructure, numpy only."""
import numpy as np


def rot_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def plane_world(seed=7, half=40.0, step=0.3, n_az=600):
    """A DEGENERATE scene for laserMapping.cpp:934-980: nothing but the ground plane z = 0.  Every plane normal is +-e_z, so
    the 3x3 normal block has one large eigenvalue and two (x, y) whose per-direction contributions stay below the
    reference's 250 / 50 thresholds -> the update must be projected onto e_z.
    Returns map [M,4] (world), scan [N,4] (lidar frame), ground-truth (R, t) and a perturbed prior (R, t)."""
    rng = np.random.default_rng(seed)
    g = np.arange(-half, half, step)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    m = np.zeros((xx.size, 4), np.float32)
    m[:, 0] = (xx.ravel() + rng.uniform(-0.1, 0.1, xx.size)).astype(np.float32)
    m[:, 1] = (yy.ravel() + rng.uniform(-0.1, 0.1, xx.size)).astype(np.float32)
    m[:, 2] = rng.normal(0.0, 0.01, xx.size).astype(np.float32)
    m[:, 3] = rng.uniform(0, 255, xx.size).astype(np.float32)
    Rgt = rot_rpy(0.01, -0.015, 0.4)
    tgt = np.array([0.5, -0.3, 1.8])
    el = np.deg2rad(np.linspace(-25.0, -3.0, 64))
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    A, E = np.meshgrid(az, el, indexing="ij")
    d_l = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    d_w = d_l @ Rgt.T
    t = -tgt[2] / d_w[:, 2]
    ok = (d_w[:, 2] < -1e-3) & (t > 0.5) & (t < 35.0)
    t = t[ok] + rng.normal(0.0, 0.01, int(ok.sum()))
    scan = np.zeros((t.size, 4), np.float32)
    scan[:, :3] = (d_l[ok] * t[:, None]).astype(np.float32)
    scan[:, 3] = rng.uniform(0, 255, t.size).astype(np.float32)
    Rp = Rgt @ rot_rpy(0.004, -0.003, 0.006)
    tp = tgt + np.array([0.08, -0.06, 0.05])
    return dict(map=m, scan=scan, Rgt=Rgt, tgt=tgt, Rprior=Rp, tprior=tp)
