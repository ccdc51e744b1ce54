"""TEST INFRASTRUCTURE: adversarial inputs for map insert + k-NN in both per-query shapes (warp / thread per query), run by
tests/test_emu_kernels.py against the SIMT emulator build.  The shapes must agree bit for bit on EVERYTHING; against the
oracle (iVox restatement) they must agree wherever the input is inside the map's documented capacity: voxel coordinates
within +-2^18, at most 7 * 128 points per voxel, finite coordinates (the rest is counted in `dropped`, include/lsdreg.h)."""
import sys

import numpy as np

import lsdreg
from oracle import oracle as O

lsdreg.init(0)
rng = np.random.default_rng(1)


def P(a):
    a = np.asarray(a, np.float32)
    out = np.zeros((a.shape[0], 4), np.float32)
    out[:, :3] = a
    return out


def case(name, pts, q, res=0.5, log2=14, vs_oracle=True, cells_equal=True):
    g = lsdreg.HashVoxelMap(res, log2)
    g.insert(pts, 7)
    o = O.OracleIvox(res, 18, 1 << 14)
    o.add(pts, 7)
    st = g.stats()
    if cells_equal:
        assert st["cells"] == o.num_cells and st["dropped"] == 0, (name, st, o.num_cells)
    for nearby in (0, 6, 18, 26, 74):
        o.set_nearby(nearby)
        for k in (1, 5):
            ref = None
            for shape in (1, 2):
                g.set_knn_shape(shape)
                r = g.knn(q, k=k, max_sq=5.0, stencil=nearby)
                if ref is None:
                    ref = r
                for a, b in zip(ref, r):
                    assert (a.view(np.int32) == b.view(np.int32)).all(), (name, nearby, k, shape)
            if vs_oracle:
                oi, od, _, oc = o.knn(q, k, 5.0)
                assert (ref[0] == oi).all() and (ref[1].view(np.int32) == od.view(np.int32)).all() and (ref[2] == oc).all(), (name, nearby, k)
    print("ok", name, st)
    sys.stdout.flush()


# points exactly on voxel borders (round half away from zero), negative coordinates, queries a hair off the border
grid = np.stack(np.meshgrid(np.arange(-6, 7), np.arange(-6, 7), np.arange(-2, 3), indexing="ij"), -1).reshape(-1, 3) * 0.25
case("borders", P(grid), P(grid + rng.choice([0, 0.25, -0.25, 1e-7, -1e-7], size=grid.shape)))
# duplicates: equal distances, ties broken by id
dup = np.repeat(rng.uniform(-3, 3, (40, 3)), 9, axis=0)
case("duplicates", P(dup), P(dup[::3]))
# a voxel beyond its capacity (> 7 * 128 points): the shapes agree; the oracle keeps what the map drops
crowd = np.concatenate([rng.uniform(-0.2, 0.2, (1200, 3)) + [10, 10, 1], rng.uniform(-4, 4, (300, 3)) + [10, 10, 1]])
case("crowded", P(crowd), P(crowd[::5]), vs_oracle=False, cells_equal=False)
# coordinates around and beyond the +-2^18 voxel limit
far = np.concatenate([rng.uniform(-2, 2, (200, 3)) + [131000.0, 0, 0], rng.uniform(-2, 2, (200, 3)) + [131072.5, -131071.0, 5],
                      rng.uniform(-2, 2, (100, 3)) + [2e5, 2e5, 0], rng.uniform(-2, 2, (200, 3))])
case("far", P(far), P(np.concatenate([far[::2], [[1e9, 0, 0], [-1e9, 1e9, 0], [131071.7, 0, 0]]])), vs_oracle=False, cells_equal=False)
# non-finite points and queries: never returned, never found
nf = rng.uniform(-3, 3, (300, 3)); nf[::17, 0] = np.nan; nf[5::23, 1] = np.inf; nf[7::29, 2] = -np.inf
qs = rng.uniform(-3, 3, (200, 3)); qs[::11, 1] = np.nan; qs[3::13, 0] = np.inf
case("nonfinite", P(nf), P(qs), cells_equal=False)
# clustered clouds, several resolutions, small tables (long probe sequences)
for t in range(3):
    n = 5000
    centres = rng.uniform(-30, 30, (30, 3))
    pts = centres[rng.integers(0, 30, n)] + rng.normal(0, [1.5, 1.5, 0.3], (n, 3))
    q = pts[rng.integers(0, n, 1000)] + rng.normal(0, 0.3, (1000, 3))
    case(f"clustered res {[0.5, 0.3, 1.0][t]}", P(pts), P(q), res=[0.5, 0.3, 1.0][t], log2=[14, 15, 13][t])
print("FUZZ_OK")
