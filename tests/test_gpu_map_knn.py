"""GPU parity: hash-voxel map insert + k-NN (K2, K3) against the CPU oracle, through the C ABI.
Bar: bit-exact neighbour ids and fp32 squared distances (integer/index work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def maps(small_world):
    import lsdreg
    from oracle import oracle as O
    m = small_world["map"]
    g = lsdreg.HashVoxelMap(0.5, 20)
    g.insert(m, 0)
    o = O.OracleIvox(0.5, 18, 1 << 18)
    o.add(m, 0)
    return g, o


def _queries(small_world):
    from oracle import oracle as O
    from lsdreg import synth
    ds = O.voxelgrid(small_world["scan"], 0.5)
    q = ds.copy()
    R, t = small_world["Rprior"], small_world["tprior"]
    q[:, :3] = (ds[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)
    return q


def test_map_stats_match_oracle(maps, small_world):
    g, o = maps
    st = g.stats()
    assert st["points"] == o.num_points == small_world["map"].shape[0]
    assert st["cells"] == o.num_cells
    assert st["dropped"] == 0


@pytest.mark.parametrize("nearby", [0, 6, 18, 26, 74])
def test_knn_stencil_bit_exact(maps, small_world, nearby):
    g, o = maps
    q = _queries(small_world)
    o.set_nearby(nearby)
    oi, od, _, oc = o.knn(q, 5, 5.0)
    gi, gd, gc = g.knn(q, 5, 5.0, nearby)
    assert (gc == oc).all()
    assert (gi == oi).all()
    assert (gd.view(np.int32) == od.view(np.int32)).all()  # bit-exact fp32 d2
    o.set_nearby(18)


@pytest.mark.parametrize("k", [1, 5, 20])
def test_knn_exact_matches_oracle(maps, small_world, k):
    import lsdreg
    g, o = maps
    q = _queries(small_world)
    oi, od, _, oc = o.knn(q, k, 5.0, exact=True)
    gi, gd, gc = g.knn(q, k, 5.0, lsdreg.STENCIL_EXACT)
    assert (gc == oc).all()
    assert (gi == oi).all()
    assert (gd.view(np.int32) == od.view(np.int32)).all()


def test_knn_empty_and_far_queries(maps):
    g, _ = maps
    idx, d2, cnt = g.knn(np.zeros((0, 4), np.float32))
    assert idx.shape == (0, 5)
    far = np.array([[1e4, 1e4, 50.0, 0.0], [-3000.0, 12.0, 3.0, 0.0]], np.float32)
    idx, d2, cnt = g.knn(far)
    assert (cnt == 0).all() and (idx == -1).all()


def test_bucket_overflow_levels():
    """> 7 points in one voxel spill to (voxel, level) lines; every point must stay findable."""
    import lsdreg
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    pts = np.zeros((500, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.2, 0.2, size=(500, 3)) + np.array([10.0, -4.0, 2.0])  # one 0.5 m voxel
    extra = np.zeros((200, 4), np.float32)
    extra[:, :3] = rng.uniform(-3, 3, size=(200, 3)) + np.array([10.0, -4.0, 2.0])
    allp = np.concatenate([pts, extra])
    g = lsdreg.HashVoxelMap(0.5, 12)
    g.insert(allp, 100)
    o = O.OracleIvox(0.5, 18, 1 << 10)
    o.add(allp, 100)
    st = g.stats()
    assert st["points"] == 700 and st["dropped"] == 0 and st["cells"] == o.num_cells
    q = allp[::7].copy()
    for k in (5, 20):
        oi, od, _, oc = o.knn(q, k, 5.0)
        gi, gd, gc = g.knn(q, k, 5.0, 18)
        assert (gi == oi).all() and (gc == oc).all()


def test_incremental_insert_equals_bulk(small_world):
    import lsdreg
    m = small_world["map"][:50000]
    a = lsdreg.HashVoxelMap(0.5, 18)
    a.insert(m, 0)
    b = lsdreg.HashVoxelMap(0.5, 18)
    for s in range(0, 50000, 7777):
        b.insert(m[s:s + 7777], s)
    assert a.stats() == b.stats()
    q = m[::50].copy()
    ia, da, ca = a.knn(q)
    ib, db, cb = b.knn(q)
    assert (ia == ib).all() and (da == db).all() and (ca == cb).all()


@pytest.mark.parametrize("nearby,k", [(18, 5), (74, 5), (18, 1), (6, 5)])
def test_batched_thread_per_query_kernel_bit_exact(maps, small_world, nearby, k):
    """From 65 536 queries on lsd_knn_query switches to the thread-per-query kernel: same bits as the oracle and as
    the warp-per-query kernel (the same queries in chunks below the threshold), overflow buckets included."""
    import lsdreg
    g, o = maps
    rng = np.random.default_rng(21)
    m = small_world["map"]
    nq = 70000
    q = m[rng.integers(0, m.shape[0], nq)].copy()
    q[:, :3] += rng.normal(0, 0.15, (nq, 3)).astype(np.float32)
    q[:100, :3] += 500.0                                            # some queries far from any voxel
    idx, d2, cnt = g.knn(q, k=k, max_sq=5.0, stencil=nearby)        # thread kernel
    parts = [g.knn(q[a:a + 30000], k=k, max_sq=5.0, stencil=nearby) for a in range(0, nq, 30000)]   # warp kernel
    np.testing.assert_array_equal(idx, np.concatenate([p[0] for p in parts]))
    np.testing.assert_array_equal(d2, np.concatenate([p[1] for p in parts]))
    np.testing.assert_array_equal(cnt, np.concatenate([p[2] for p in parts]))
    o.set_nearby(nearby)
    sub = rng.integers(0, nq, 4000)
    oi, od, _, oc = o.knn(q[sub], k, 5.0)
    np.testing.assert_array_equal(idx[sub], oi)
    np.testing.assert_array_equal(d2[sub].view(np.int32), od.view(np.int32))
    np.testing.assert_array_equal(cnt[sub], oc)
    o.set_nearby(18)
    # a map with 40-point buckets: the overflow-level walk of the thread kernel
    gm = lsdreg.HashVoxelMap(0.5, 14)
    pts = np.zeros((4000, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.2, 0.2, (4000, 3)) + rng.integers(0, 10, (4000, 1)) * np.array([[0.5, 0, 0]])
    gm.insert(pts, 0)
    qq = np.repeat(pts[:700], 100, axis=0)[:70000].copy()
    a = gm.knn(qq, k=k, max_sq=5.0, stencil=nearby)
    b = [gm.knn(qq[s:s + 30000], k=k, max_sq=5.0, stencil=nearby) for s in range(0, 70000, 30000)]
    for j in range(3):
        np.testing.assert_array_equal(a[j], np.concatenate([p[j] for p in b]))


def test_box_delete_matches_port():
    """lsd_map_delete_boxes (KD_TREE::Delete_Point_Boxes semantics) against the CPU port, itself pinned against the
    compiled ikd-Tree (tests/test_oracle_golden.py): same count, and no query — stencil or exact, either kernel shape —
    ever returns a deleted point."""
    import lsdreg
    from oracle import oracle as O
    rng = np.random.default_rng(12)
    pts = np.zeros((60000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (60000, 3)); pts[:, 2] *= 0.1          # ~9 points per 0.5 m voxel: overflow lines too
    boxes = np.array([[-5, -5, -2, 3, 4, 2], [8, -20, -4, 20, -10, 4], [100, 100, 100, 101, 101, 101]], np.float32)
    boxes[0, 0] = pts[7, 0]; boxes[0, 3] = pts[9, 0]
    g = lsdreg.HashVoxelMap(0.5, 16); g.insert(pts, 0)
    o = O.OracleIvox(0.5, 18, 1 << 16); o.add(pts, 0)
    n_g, n_o = g.delete_boxes(boxes), o.delete_boxes(boxes)
    assert n_g == n_o > 3000
    assert g.stats()["points"] == o.num_points == len(pts) - n_o
    assert g.delete_boxes(boxes) == 0                                        # idempotent
    q = pts[rng.integers(0, len(pts), 70000)].copy(); q[:, :3] += rng.normal(0, 0.2, (70000, 3)).astype(np.float32)
    oi, od, _, oc = o.knn(q[:5000], 5, 5.0)
    for sl in (slice(0, 5000), slice(0, 70000)):                             # warp-per-query and thread-per-query kernels
        gi, gd, gc = g.knn(q[sl], 5, 5.0, 18)
        np.testing.assert_array_equal(gi[:5000], oi); np.testing.assert_array_equal(gc[:5000], oc)
        np.testing.assert_array_equal(gd[:5000].view(np.int32), od.view(np.int32))
    oi, od, _, oc = o.knn(q[:3000], 20, 5.0, exact=True)
    gi, gd, gc = g.knn(q[:3000], 20, 5.0, lsdreg.STENCIL_EXACT)
    np.testing.assert_array_equal(gi, oi); np.testing.assert_array_equal(gc, oc)
    inside = (pts[:, 0] >= boxes[0, 0]) & (pts[:, 0] < boxes[0, 3]) & (pts[:, 1] >= -5) & (pts[:, 1] < 4) & (pts[:, 2] >= -2) & (pts[:, 2] < 2)
    assert not np.isin(gi[gi >= 0], np.nonzero(inside)[0]).any()
    more = pts[:1000].copy(); more[:, 0] += 0.01                              # the map keeps working after a delete
    g.insert(more, 100000); o.add(more, 100000)
    gi, gd, gc = g.knn(q[:2000], 5, 5.0, 18); oi, od, _, oc = o.knn(q[:2000], 5, 5.0)
    np.testing.assert_array_equal(gi, oi); np.testing.assert_array_equal(gc, oc)
