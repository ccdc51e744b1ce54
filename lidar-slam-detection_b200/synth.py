"""Seeded synthetic inputs for the registration hot path.

The reference ships no datasets, fixtures or golden vectors, so every test and
benchmark input is generated here, deterministically, with numpy only:

* ``scan64``     a 64-beam spinning-lidar scan ray-cast into a procedural block scene
* ``block_map``  a map sampled on the same surfaces, one or many tiled blocks (2 km x 2 km -> ~10 M)
* ``perturb``    small seeded SE(3) offsets for pose priors

Scene ("block"): ground plane z=0, the 4 inner walls of a BX x BY box, NBOX axis-aligned boxes
with seeded positions/sizes.  Blocks tile the plane with period (BX, BY); each block draws its own
boxes from ``seed + block index`` so different blocks are different maps.
"""
from __future__ import annotations

import numpy as np

BX, BY, WALL_H = 120.0, 80.0, 8.0
NBOX = 40
SEED0 = 20260922


def block_boxes(seed: int, bi: int, bj: int) -> np.ndarray:
    """[NBOX, 6] = (xmin, ymin, zmin, xmax, ymax, zmax) in WORLD coordinates for block (bi, bj)."""
    rng = np.random.default_rng([seed, bi + 1000, bj + 1000])
    size = rng.uniform(1.0, 6.0, size=(NBOX, 3))
    size[:, 2] = rng.uniform(1.0, 5.0, size=NBOX)
    cx = rng.uniform(8.0, BX - 8.0, size=NBOX)
    cy = rng.uniform(8.0, BY - 8.0, size=NBOX)
    # keep the block centre (sensor spawn area) free
    near = (np.abs(cx - BX / 2) < 6.0) & (np.abs(cy - BY / 2) < 6.0)
    cx[near] += 14.0
    lo = np.stack([cx - size[:, 0] / 2, cy - size[:, 1] / 2, np.zeros(NBOX)], 1)
    hi = np.stack([cx + size[:, 0] / 2, cy + size[:, 1] / 2, size[:, 2]], 1)
    off = np.array([bi * BX, bj * BY, 0.0])
    return np.concatenate([lo + off, hi + off], 1)


def block_center(bi: int, bj: int, z: float = 1.8) -> np.ndarray:
    return np.array([bi * BX + BX / 2, bj * BY + BY / 2, z])


def rot_from_rpy(r: float, p: float, y: float) -> np.ndarray:
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def perturb(seed: int, max_t: float = 0.3, max_deg: float = 1.0):
    """Seeded small SE(3): returns (R [3,3], t [3])."""
    rng = np.random.default_rng([seed, 7])
    ang = np.deg2rad(rng.uniform(-max_deg, max_deg, size=3))
    t = rng.uniform(-max_t, max_t, size=3)
    return rot_from_rpy(*ang), t


def scan64(seed: int, n_az: int = 1563, R: np.ndarray | None = None, t: np.ndarray | None = None,
           bi: int = 0, bj: int = 0, noise: float = 0.02, n_beams: int = 64) -> np.ndarray:
    """Ray-cast a spinning lidar at world pose (R, t) inside block (bi, bj).

    Returns float32 [N, 4] = (x, y, z, intensity) in the LIDAR frame, N <= n_beams * n_az (rays
    that hit nothing within 200 m are dropped).  Order: azimuth-major, beam-minor (firing order).
    n_az=1563 -> ~100 k points (config 2), n_az=250 -> 16 k (config 1).
    """
    rng = np.random.default_rng([seed, 11])
    if R is None:
        R = np.eye(3)
    if t is None:
        t = block_center(bi, bj)
    elev = np.deg2rad(np.linspace(-25.0, 15.0, n_beams))
    az = np.linspace(0.0, 2 * np.pi, n_az, endpoint=False)
    A, E = np.meshgrid(az, elev, indexing="ij")
    d_l = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    d = d_l @ R.T
    o = np.asarray(t, dtype=np.float64)
    rmax = 200.0
    best = np.full(d.shape[0], rmax)
    with np.errstate(divide="ignore", invalid="ignore"):
        # ground z = 0
        tg = -o[2] / d[:, 2]
        tg = np.where((d[:, 2] < 0) & (tg > 0), tg, rmax)
        best = np.minimum(best, tg)
        # inner walls of the block (planes x = x0|x1, y = y0|y1, height WALL_H)
        x0, y0 = bi * BX, bj * BY
        for axis, val in ((0, x0), (0, x0 + BX), (1, y0), (1, y0 + BY)):
            tw = (val - o[axis]) / d[:, axis]
            hz = o[2] + tw * d[:, 2]
            ok = (tw > 0) & (hz >= 0) & (hz <= WALL_H)
            best = np.minimum(best, np.where(ok, tw, rmax))
        # boxes: slab test
        for b in block_boxes(SEED0, bi, bj):
            t1 = (b[:3] - o) / d
            t2 = (b[3:] - o) / d
            tn = np.minimum(t1, t2).max(1)
            tf = np.maximum(t1, t2).min(1)
            ok = (tn <= tf) & (tn > 0)
            best = np.minimum(best, np.where(ok, tn, rmax))
    hit = best < rmax
    rr = best + rng.normal(0.0, noise, size=best.shape)
    pts = d_l * rr[:, None]
    inten = rng.uniform(0.0, 255.0, size=best.shape)
    out = np.concatenate([pts, inten[:, None]], 1)[hit]
    return np.ascontiguousarray(out, dtype=np.float32)


def _sample_block(rng: np.random.Generator, bi: int, bj: int, spacing: float) -> np.ndarray:
    """Jittered-grid surface samples of one block at ~`spacing` metres (world coordinates)."""
    x0, y0 = bi * BX, bj * BY
    parts = []
    # ground
    gx = np.arange(x0 + spacing / 2, x0 + BX, spacing)
    gy = np.arange(y0 + spacing / 2, y0 + BY, spacing)
    X, Y = np.meshgrid(gx, gy, indexing="ij")
    g = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], 1)
    g[:, :2] += rng.uniform(-0.4, 0.4, size=(g.shape[0], 2)) * spacing
    parts.append(g)
    # walls
    wz = np.arange(spacing / 2, WALL_H, spacing)
    for axis, val in ((0, x0), (0, x0 + BX), (1, y0), (1, y0 + BY)):
        along = gy if axis == 0 else gx
        U, Z = np.meshgrid(along, wz, indexing="ij")
        w = np.zeros((U.size, 3))
        w[:, axis] = val
        w[:, 1 - axis] = U.ravel() + rng.uniform(-0.4, 0.4, size=U.size) * spacing
        w[:, 2] = Z.ravel() + rng.uniform(-0.4, 0.4, size=U.size) * spacing
        parts.append(w)
    # boxes: 4 sides + top
    for b in block_boxes(SEED0, bi, bj):
        lo, hi = b[:3], b[3:]
        for axis in (0, 1):
            oth = 1 - axis
            u = np.arange(lo[oth] + spacing / 2, hi[oth], spacing)
            z = np.arange(spacing / 2, hi[2], spacing)
            if u.size == 0 or z.size == 0:
                continue
            U, Z = np.meshgrid(u, z, indexing="ij")
            for val in (lo[axis], hi[axis]):
                s = np.zeros((U.size, 3))
                s[:, axis] = val
                s[:, oth] = U.ravel() + rng.uniform(-0.3, 0.3, size=U.size) * spacing
                s[:, 2] = Z.ravel() + rng.uniform(-0.3, 0.3, size=U.size) * spacing
                parts.append(s)
        u = np.arange(lo[0] + spacing / 2, hi[0], spacing)
        v = np.arange(lo[1] + spacing / 2, hi[1], spacing)
        if u.size and v.size:
            U, V = np.meshgrid(u, v, indexing="ij")
            s = np.stack([U.ravel(), V.ravel(), np.full(U.size, hi[2])], 1)
            parts.append(s)
    return np.concatenate(parts, 0)


def block_map(seed: int, blocks_x: int = 1, blocks_y: int = 1, spacing: float = 0.5,
              noise: float = 0.01) -> np.ndarray:
    """Surface-sampled map of blocks_x x blocks_y tiled blocks, float32 [N, 4] (x,y,z,intensity).

    One point per ~spacing^2 of surface (what the reference's 0.5 m keep-closest map thinning
    converges to, laserMapping.cpp:540-560).  1x1 @0.5 -> ~49 k; 2x2 -> ~200 k (config 1);
    16x25 @0.62 -> ~10 M (config 2).
    """
    rng = np.random.default_rng([seed, 13])
    parts = []
    for bi in range(blocks_x):
        for bj in range(blocks_y):
            parts.append(_sample_block(rng, bi, bj, spacing))
    p = np.concatenate(parts, 0)
    p += rng.normal(0.0, noise, size=p.shape)
    inten = rng.uniform(0.0, 255.0, size=(p.shape[0], 1))
    return np.ascontiguousarray(np.concatenate([p, inten], 1), dtype=np.float32)
