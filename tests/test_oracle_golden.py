"""CPU: the plain-C oracle port against (a) golden vectors generated from the COMPILED reference
(tests/golden/make_golden.py) and (b) the compiled reference itself when oracle/_ref is present.
Bars: neighbour id sets bit-exact; fp32 k-d tree distances bit-exact; plane residual |pd2| within 2e-4 m
(Eigen's vectorised QR reductions are not bit-reproducible by scalar code, see oracle/lsd_oracle.c)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lio_ref_small.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.fixture(scope="module")
def port_map(gold):
    iv = O.OracleIvox(0.5, 18, 1 << 16)
    iv.add(gold["map"], 0)
    return iv


@pytest.mark.parametrize("nearby", [18, 74])
def test_ivox_knn_matches_reference_golden(gold, port_map, nearby):
    port_map.set_nearby(nearby)
    ids, d2, xyz, cnt = port_map.knn(gold["query"], 5, 5.0)
    port_map.set_nearby(18)
    assert (cnt == gold[f"ivox{nearby}_cnt"]).all()
    assert (np.sort(ids, 1) == gold[f"ivox{nearby}_ids"]).all()


def test_exact_knn_matches_ikdtree_golden(gold, port_map):
    ids, d2, xyz, cnt = port_map.knn(gold["query"], 5, 5.0, exact=True)
    kd_ok = (gold["ikd_cnt"] >= 5) & (gold["ikd_d2"][:, 4] <= 5)       # laserMapping.cpp:846-847
    assert ((cnt >= 5) == kd_ok).all()
    assert (np.sort(ids[kd_ok], 1) == np.sort(gold["ikd_ids"][kd_ok], 1)).all()
    assert (d2[kd_ok].view(np.int32) == gold["ikd_d2"][kd_ok].view(np.int32)).all()  # same fp32 arithmetic


def test_esti_plane_matches_reference_golden(gold):
    pa, ok = O.esti_plane(gold["plane_in"], 0.1)
    ra, rok = gold["plane_abcd"], gold["plane_ok"]
    pts = gold["plane_in"]
    cen = pts.mean(1)
    pd = (pa[:, :3] * cen).sum(1) + pa[:, 3]
    rpd = (ra[:, :3] * cen).sum(1) + ra[:, 3]
    # accept/reject may only differ for planes within rounding of the 0.1 m threshold
    flip = ok != rok
    assert flip.mean() < 0.01
    good = (ok == 1) & (rok == 1)
    assert np.abs(pd - rpd)[good].max() < 2e-4
    assert np.abs((pa[:, :3] * ra[:, :3]).sum(1))[good].min() > 1 - 1e-5  # same normal up to sign (same sign by construction)


def _plane_cloud(rng, n, far, noise):
    c = rng.uniform(-far, far, (n, 1, 3)); c[:, :, 2] = rng.uniform(-5, 30, (n, 1))
    nrm = rng.normal(size=(n, 1, 3)); nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    p = rng.uniform(-0.7, 0.7, (n, 5, 3))
    p = p - (p * nrm).sum(2, keepdims=True) * nrm + nrm * rng.normal(scale=noise, size=(n, 5, 1))
    return np.ascontiguousarray((c + p).astype(np.float32))


@pytest.mark.skipif(not (O.HAVE_REF and hasattr(O.ref, "ref_esti_plane_qr")), reason="oracle/_ref not built (needs /root/reference)")
def test_esti_plane_bit_exact_vs_compiled_reference():
    """esti_plane<float> (common_lib.h:236-268) compiled unmodified, Eigen's ColPivHouseholderQR in the x86-64 build the reference
    arm runs, against the plain-C restatement: EVERY bit of the plane coefficients, the accept flag, and of every intermediate of
    the factorisation (packed QR, Householder coefficients, permutation, non-zero pivots, solution) — near the origin, 1.5 km
    from it (where the fp32 solve is ill-conditioned and any change of summation order shows), on noisy neighbourhoods the gate
    rejects, and on rank-deficient inputs (all-zero, coplanar through the origin, five copies of one point)."""
    rng = np.random.default_rng(20260923)
    sets = [_plane_cloud(rng, 40000, far, noise) for far, noise in ((1500.0, 0.01), (1500.0, 0.2), (50.0, 0.01), (1.0, 0.05))]
    deg = _plane_cloud(rng, 3000, 20.0, 0.0)
    deg[:1000] = 0; deg[1000:2000, :, 2] = 0; deg[2000:] = deg[2000:, :1]
    sets.append(deg)
    for pts in sets:
        a, oka = O.esti_plane(pts, 0.1)
        b, okb = O.ref_esti_plane(pts, 0.1)
        assert (oka == okb).all()
        assert (a.view(np.uint32) == b.view(np.uint32)).all()      # NaN patterns of the degenerate inputs included
        qa, qb = O.esti_plane_qr(pts), O.ref_esti_plane_qr(pts)
        for k in ("qr", "hcoeffs", "x"):
            assert (qa[k].view(np.uint32) == qb[k].view(np.uint32)).all(), k
        assert (qa["perm"] == qb["perm"]).all() and (qa["nonzero_pivots"] == qb["nonzero_pivots"]).all()


def _nth_sequences(rng, trials):
    for t in range(trials):
        n = int(rng.integers(1, 257)) if t % 3 else int(rng.integers(1, 33))
        kind = t % 5
        if kind == 0: d = rng.random(n)
        elif kind == 1: d = np.round(rng.random(n) * 6) / 6                       # many ties
        elif kind == 2: d = np.sort(rng.random(n))
        elif kind == 3: d = np.sort(rng.random(n))[::-1].copy()
        else: d = np.concatenate([np.arange(n // 2)[::-1], np.arange(n - n // 2)]).astype(float)   # organ pipe
        if t % 2:
            first = int(rng.integers(0, n)); last = int(rng.integers(first, n + 1)); nth = int(rng.integers(first, last + 1))
        else:
            first, last = 0, n; nth = min(4, n) if t % 4 == 0 else 0
        yield np.ascontiguousarray(d, np.float32), first, min(nth, last), last


def _depth_limit_sequences():
    rows = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nth_element_depth_limit_sequences.npy"))
    for r in rows:
        n, nth = int(r[0]), int(r[1])
        yield np.ascontiguousarray(r[2:2 + n], np.float32), 0, nth, n


@pytest.mark.skipif(not (O.HAVE_REF and hasattr(O.ref, "ref_std_nth_element")), reason="oracle/_ref not built (needs /root/reference)")
def test_restated_nth_element_is_std_nth_element():
    """oracle/lsd_oracle.c::ref_nth_element (libstdc++'s __introselect restated) against std::nth_element itself, called on a
    std::vector of the reference's own DistPoint exactly as IVox::GetClosestPoint calls it: the same permutation on random,
    tied, sorted, reversed and organ-pipe sequences of 1-256 elements, on sub-ranges, and on the committed adversarial
    sequences that drive introselect past its depth limit into __heap_select (checked: the restatement's counter moves)."""
    import ctypes as C
    O.port.orc_heap_select_count.restype = C.c_long

    def both(d, first, nth, last):
        n = d.shape[0]
        d64 = d.astype(np.float64)
        a = np.arange(n, dtype=np.int32); b = a.copy()
        O.port.orc_nth_element(C.c_void_p(d.ctypes.data), C.c_void_p(a.ctypes.data), n, first, nth, last)
        O.ref.ref_std_nth_element(C.c_void_p(d64.ctypes.data), C.c_void_p(b.ctypes.data), n, first, nth, last)
        return a, b
    for d, first, nth, last in _nth_sequences(np.random.default_rng(3), 6000):
        a, b = both(d, first, nth, last)
        assert (a == b).all(), (d.shape[0], first, nth, last)
    before = O.port.orc_heap_select_count()
    k = 0
    for d, first, nth, last in _depth_limit_sequences():
        a, b = both(d, first, nth, last)
        assert (a == b).all()
        k += 1
    assert k >= 8 and O.port.orc_heap_select_count() - before == k


@pytest.mark.skipif(not O.HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("npts,spread,mode,nearby", [(20000, 6.0, "rand", 18), (200000, 6.0, "rand", 18), (20000, 6.0, "dup", 18),
                                                     (60000, 4.0, "lattice", 18), (100000, 5.0, "rand", 74), (100000, 5.0, "rand", 26),
                                                     (3000, 6.0, "rand", 18), (30000, 5.0, "rand", 6)])
def test_reference_neighbour_order_vs_compiled_ivox(npts, spread, mode, nearby):
    """IVox::GetClosestPoint returns its neighbours in the order two std::nth_element calls (plus one per crowded voxel) leave
    them in (ivox3d.h:159-164, ivox3d_node.hpp:118-123), and esti_plane's fp32 solve depends on that order.  The restated
    introselect (oracle/lsd_oracle.c::ref_nth_element, libstdc++'s algorithm) on the restated candidate sequence gives the
    compiled reference's rows id for id, position for position: sparse and crowded voxels (14 points per voxel: the per-voxel
    truncation), fewer than k in range, exact duplicates and lattice points (distance ties), every stencil."""
    rng = np.random.default_rng([7, npts, nearby])
    pts = np.zeros((npts, 4), np.float32); pts[:, :3] = rng.uniform(-spread, spread, (npts, 3))
    q = np.zeros((15000, 4), np.float32); q[:, :3] = rng.uniform(-spread, spread, (15000, 3))
    if mode == "dup":
        pts[npts // 2:, :3] = pts[:npts - npts // 2, :3]
    if mode == "lattice":
        pts[:, :3] = np.round(pts[:, :3] * 8) / 8; q[:, :3] = np.round(q[:, :3] * 4) / 4
    po, rf = O.OracleIvox(0.5, nearby, 1 << 16), O.RefIvox(0.5, nearby)
    step = npts // 3 + 1
    for a in range(0, npts, step):
        po.add(pts[a:a + step], a); rf.add(pts[a:a + step], a)
    ids, d2, xyz, cnt = po.knn(q, 5, 5.0, reference_order=True)
    rid, rxyz, rcnt = rf.knn(q, 5, 5.0)
    assert (cnt == rcnt).all()
    assert (ids == rid).all()
    assert (xyz.view(np.uint32) == rxyz.view(np.uint32)).all()
    cid, cd2, _, _ = po.knn(q, 5, 5.0)                       # the canonical order holds the same distances, and the same
    assert (np.sort(cd2, 1) == np.sort(d2, 1)).all()         # ids unless two points tie for the fifth place (nth_element's call)
    if mode == "rand":
        assert (np.sort(cid, 1) == np.sort(ids, 1)).all()


@pytest.mark.skipif(not O.HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_compiled_reference_on_fresh_data():
    from lsdreg import synth
    m = synth.block_map(21, 1, 2, 0.5)
    scan = synth.scan64(22, 120)
    q = O.voxelgrid(scan, 0.5)
    q[:, :3] += synth.block_center(0, 0).astype(np.float32)
    iv = O.OracleIvox(0.5, 18, 1 << 17); iv.add(m, 0)
    rv = O.RefIvox(0.5, 18); rv.add(m, 0)
    assert iv.num_cells == rv.num_cells
    ids, d2, xyz, cnt = iv.knn(q)
    rids, rxyz, rcnt = rv.knn(q)
    assert (cnt == rcnt).all() and (np.sort(ids, 1) == np.sort(rids, 1)).all()
    # reference order: nearest neighbour first (ivox3d.h:163), ours: canonical ascending
    has = cnt > 0
    assert (ids[has, 0] == rids[has, 0]).all() or (d2[has, 0] == d2[has, 1]).any()
    kd = O.RefIkd(); kd.build(m)
    kids, kd2, kcnt = kd.knn(q)
    eids, ed2, _, ecnt = iv.knn(q, 5, 5.0, exact=True)
    acc = (kcnt >= 5) & (kd2[:, 4] <= 5)
    assert ((ecnt >= 5) == acc).all()
    assert (np.sort(kids[acc], 1) == np.sort(eids[acc], 1)).all() and (kd2[acc] == ed2[acc]).all()


def test_voxelgrid_properties():
    """PCL VoxelGrid restatement (parity unpinned: PCL is not vendored): structural properties."""
    from lsdreg import synth
    scan = synth.scan64(5, 200)
    out, vidx = O.voxelgrid(scan, 0.5, want_vidx=True)
    assert (np.diff(vidx) > 0).all()                       # ascending, unique leaf indices
    inv = np.float32(1.0) / np.float32(0.5)
    leaf_in = np.floor(scan[:, :3] * inv).astype(np.int64)
    leaf_out = np.floor(out[:, :3] * inv).astype(np.int64)
    assert len(np.unique(leaf_in, axis=0)) == out.shape[0]  # one output per occupied leaf
    # centroid lies in (or on the boundary of) its leaf
    key_in = set(map(tuple, leaf_in))
    near = [tuple(k) in key_in for k in leaf_out]
    assert np.mean(near) > 0.999
    # idempotent up to rounding: filtering the centroids keeps one point per leaf
    again = O.voxelgrid(out, 0.5)
    assert abs(again.shape[0] - out.shape[0]) <= 0.001 * out.shape[0] + 2
    assert O.voxelgrid(np.zeros((0, 4), np.float32), 0.5).shape == (0, 4)


def test_oracle_lio_converges():
    from lsdreg import synth
    from oracle import eskf
    from oracle.lio import OracleLio
    m = synth.block_map(1, 1, 1, 0.5)
    lio = OracleLio(18, expected_cells=1 << 17)
    lio.add_map_points(m)
    Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
    tgt = synth.block_center(0, 0) + np.array([1.0, -2.0, 0.0])
    scan = synth.scan64(2, 120, Rgt, tgt)
    dR, dt = synth.perturb(5)
    prior = eskf.State(); prior.rot = eskf.R_to_quat(Rgt @ dR); prior.pos = tgt + dt
    r = lio.process_scan(scan, prior, eskf.init_P())
    assert r["iters"] >= 3 and r["log"][-1]["n_eff"] > 500
    assert np.abs(lio.x.pos - tgt).max() < 0.03
    assert np.linalg.norm(eskf.so3_log(eskf.R_to_quat(Rgt.T @ eskf.quat_to_R(lio.x.rot)))) < 2e-3


@pytest.mark.skipif(not (O.HAVE_REF and hasattr(O.ref, "ref_lio_hmodel")), reason="oracle/_ref not built")
def test_full_lio_loop_port_equals_reference_classes():
    """The whole per-scan loop (k-NN, plane fit, gate, H rows, degeneracy, ESKF, map_incremental) run on
    the COMPILED reference IVox + esti_plane (Eigen QR) + Eigen eigensolver vs the plain-C port:
    identical effective-point counts at every iteration, identical insert decisions, pose within
    1e-6 m / 1e-7 rad (the only difference left is QR rounding)."""
    from lsdreg import synth
    from oracle import eskf
    from oracle.lio import OracleLio
    m = synth.block_map(1, 1, 2, 0.5)
    Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
    tgt = synth.block_center(0, 0) + np.array([1.0, -2.0, 0.0])
    scan = synth.scan64(2, 200, Rgt, tgt)
    dR, dt = synth.perturb(5)
    out = {}
    for be in ("port", "reference"):
        lio = OracleLio(18, expected_cells=1 << 17, backend=be)
        lio.add_map_points(m)
        prior = eskf.State(); prior.rot = eskf.R_to_quat(Rgt @ dR); prior.pos = tgt + dt
        r = lio.process_scan(scan, prior, eskf.init_P())
        out[be] = (lio.x.to_vec(), [l["n_eff"] for l in r["log"]], r["added"], lio.map.num_cells)
    (xp, np_, ap, cp), (xr, nr, ar, cr) = out["port"], out["reference"]
    assert np_ == nr and ap == ar and cp == cr
    assert np.abs(xp[:3] - xr[:3]).max() < 1e-6
    assert np.linalg.norm(eskf.so3_log(eskf.quat_mul(eskf.quat_conj(xp[3:7]), xr[3:7]))) < 1e-7


def test_box_delete_port_equals_compiled_ikdtree():
    """KD_TREE::Delete_Point_Boxes on the compiled reference tree vs the port's hash map: same count, and the exact k-NN
    of both after the delete are identical (ids and distances)."""
    if not O.HAVE_REF or not hasattr(O.ref, "ref_ikd_delete_boxes"):
        pytest.skip("oracle/_ref/libref_lio.so without ref_ikd_delete_boxes")
    rng = np.random.default_rng(12)
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (20000, 3)); pts[:, 2] *= 0.2
    boxes = np.array([[-5, -5, -2, 3, 4, 2], [8, -20, -4, 20, -10, 4], [100, 100, 100, 101, 101, 101]], np.float32)
    boxes[0, 0] = pts[7, 0]; boxes[0, 3] = pts[9, 0]      # a point exactly on min (deleted) and one exactly on max (kept)
    tree = O.RefIkd(); tree.build(pts, 0)
    m = O.OracleIvox(0.5, 18, 1 << 16); m.add(pts, 0)
    n_ref = tree.delete_boxes(boxes)
    n_port = m.delete_boxes(boxes)
    assert n_ref == n_port > 1000
    assert m.num_points == len(pts) - n_port
    q = pts[rng.integers(0, len(pts), 3000)].copy(); q[:, :3] += rng.normal(0, 0.2, (3000, 3)).astype(np.float32)
    ri, rd, rc = tree.knn(q, 5)
    pi, pd, _, pc = m.knn(q, 5, 1e9, exact=True)
    ri, rd = O.canonical_rows(ri, rd)
    np.testing.assert_array_equal(rc, pc)
    np.testing.assert_array_equal(ri, pi)
    np.testing.assert_array_equal(rd.view(np.int32), pd.view(np.int32))
