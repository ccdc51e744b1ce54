"""GPU parity: the brick layout of the map and the TMA-staged batched k-NN over it (csrc/brick.cuh, lsd_knn_set_shape 3).
Bar: ids, fp32 d2 and counts BIT-IDENTICAL to the line-based kernels (shapes 1 / 2) and to the plain-C port of
IVox::GetClosestPoint (ivox3d.h:139-171; the port is pinned bit-exact to the compiled iVox by tests/test_oracle_golden.py),
for every stencil the pages serve (CENTER / NEARBY6 / 18 / 26) and k in {1, 5} — whether the pages were filled by the
inserts themselves (dual write), copied from a populated map (lsd_map_enable_bricks after the fact), grown by later
inserts, or thinned by lsd_map_delete_boxes; through crowded bricks (several pages per brick), queries far from the map,
on brick borders, at NaN, and beyond the +-2^18 voxel range."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STENCILS = (0, 6, 18, 26)


def _queries(m, n, seed, sigma=0.15):
    rng = np.random.default_rng(seed)
    q = m[rng.integers(0, m.shape[0], n)].copy()
    q[:, :3] += rng.normal(0.0, sigma, (n, 3)).astype(np.float32)
    return np.ascontiguousarray(q, np.float32)


def _same(a, b, what):
    for j, (u, v) in enumerate(zip(a, b)):
        assert (u.view(np.int32) == v.view(np.int32)).all(), (what, j, int((u.view(np.int32) != v.view(np.int32)).sum()))


def _check_all(g, o, q, what, vs_oracle=True):
    for nearby in STENCILS:
        if o is not None:
            o.set_nearby(nearby)
        for k in (1, 5):
            g.set_knn_shape(2)
            ref = g.knn(q, k=k, max_sq=5.0, stencil=nearby)
            g.set_knn_shape(3)
            got = g.knn(q, k=k, max_sq=5.0, stencil=nearby)
            _same(got, ref, f"{what}: bricks vs thread shape, stencil {nearby} k {k}")
            if o is not None and vs_oracle:
                oi, od, _, oc = o.knn(q, k, 5.0)
                assert (got[0] == oi).all() and (got[1].view(np.int32) == od.view(np.int32)).all() and (got[2] == oc).all(), (what, nearby, k)
    g.set_knn_shape(0)


def test_brick_pages_bit_identical(small_world):
    import lsdreg
    from oracle import oracle as O
    m = small_world["map"]
    q = np.concatenate([_queries(m, 6000, 3), _queries(m, 1500, 4, 1.5),
                        np.array([[5000.0, 5000.0, 0, 0], [np.nan, 1, 1, 0], [1, np.inf, 1, 0], [2.0e5, 0, 0, 0], [-131072.2, 3, 1, 0]], np.float32)])
    o = O.OracleIvox(0.5, 18, 1 << 18)
    o.add(np.ascontiguousarray(m[:, :3]), 0)
    # (a) dual write: pages filled by the inserts themselves
    g = lsdreg.HashVoxelMap(0.5, 20)
    g.enable_bricks(15)
    g.insert(m, 0)
    st = g.brick_stats()
    assert st["dropped"] == 0 and st["replicas"] >= g.stats()["points"] and st["pages"] > 100, st
    _check_all(g, o, q, "dual write")
    # the same batch in spatial order (every bin counter contended, long runs of one brick)
    cell = np.round(q[:, :3] / 0.5)
    cell[~np.isfinite(cell)] = 0
    _check_all(g, o, np.ascontiguousarray(q[np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))]), "ordered batch")
    # (b) pages copied from a populated map, then grown by a second insert
    h = lsdreg.HashVoxelMap(0.5, 20)
    half = m.shape[0] // 2
    h.insert(m[:half], 0)
    h.enable_bricks(15)
    h.insert(m[half:], half)
    assert h.brick_stats() == st
    _check_all(h, o, q, "rebuild + grow")
    # (c) box delete: replicas get the same tombstones
    boxes = np.array([[50, 30, -1, 70, 50, 3], [0, 0, 0.5, 200, 200, 2.0]], np.float32)
    assert g.delete_boxes(boxes) == o.delete_boxes(boxes)
    _check_all(g, o, q, "after delete_boxes")
    # shape 3 serves reach-1 stencils only: asked for explicitly on NEARBY74 it is an error, in auto mode the lines answer
    g.set_knn_shape(3)
    with pytest.raises(lsdreg.LsdError):
        g.knn(q[:10], k=5, stencil=74)
    g.set_knn_shape(0)
    g.knn(q[:10], k=5, stencil=74)
    plain = lsdreg.HashVoxelMap(0.5, 16)
    plain.set_knn_shape(3)
    with pytest.raises(lsdreg.LsdError):
        plain.knn(q[:10], k=5)


def test_crowded_bricks_and_borders():
    """More than 232 points in one brick region (level pages), every point and query on voxel / brick borders."""
    import lsdreg
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    crowd = np.concatenate([rng.uniform(-1.9, 1.9, (3000, 3)) * [1, 1, 0.4] + [10, 10, 0.3],      # ~3000 points inside one brick
                            rng.uniform(-6, 6, (2000, 3)) + [10, 10, 0.5]])
    grid = np.stack(np.meshgrid(np.arange(-20, 21), np.arange(-20, 21), np.arange(-6, 7), indexing="ij"), -1).reshape(-1, 3) * 0.25
    pts = np.zeros((crowd.shape[0] + grid.shape[0], 4), np.float32)
    pts[:, :3] = np.concatenate([crowd, grid])
    q = pts[::3].copy()
    q[:, :3] += rng.choice([0, 0.25, -0.25, 1e-6, -1e-6], size=(q.shape[0], 3)).astype(np.float32)
    g = lsdreg.HashVoxelMap(0.5, 18)
    g.enable_bricks(12)
    g.insert(pts, 100)
    assert g.brick_stats()["dropped"] == 0 and g.stats()["dropped"] == 0
    o = O.OracleIvox(0.5, 18, 1 << 16)
    o.add(np.ascontiguousarray(pts[:, :3]), 100)
    _check_all(g, o, q, "crowded + borders")


def test_auto_shape_uses_the_pages_for_large_batches(small_world):
    import lsdreg
    m = small_world["map"]
    q = _queries(m, 70000, 9)
    g = lsdreg.HashVoxelMap(0.5, 20)
    g.insert(m, 0)
    g.set_knn_shape(2)
    ref = g.knn(q, k=5)
    g.set_knn_shape(0)
    g.enable_bricks(15)
    got = g.knn(q, k=5)          # >= 65 536 queries: auto = brick pages
    _same(got, ref, "auto shape")
