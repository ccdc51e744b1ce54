"""CPU: the ScanContext restatement (oracle/scancontext.py, row N4) pinned against the reference's own Scancontext.cpp,
compiled unmodified into oracle/_ref/libref_keyframe.so (oracle/ref_keyframe.cpp).  Synthetic, seeded: 64-beam scans of the
procedural scene seen from poses spread over several blocks, so the database holds revisits (same place, other yaw),
near misses and unrelated places.  Bars: descriptor bit-exact; keys, distances bit-exact (the restatement reproduces
Eigen's SSE2 reduction order); shifts, candidate sets, closest match and candidate list identical."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import scancontext as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "libref_keyframe.so")
pytestmark = pytest.mark.skipif(not os.path.exists(_REF), reason="oracle/_ref/libref_keyframe.so not built (needs /root/reference)")

_pd, _pf, _pi = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(_REF)
    L.ref_sc_make.argtypes = [_pf, C.c_int, C.c_double, C.c_double, _pd, _pd, _pd]
    L.ref_sc_make.restype = None
    L.ref_sc_distance.argtypes = [_pd, _pd, _pd, _pi]
    L.ref_sc_distance.restype = None
    L.ref_sc_db_create.argtypes = [_pd, C.c_int, C.c_double]
    L.ref_sc_db_create.restype = C.c_void_p
    L.ref_sc_db_destroy.argtypes = [C.c_void_p]
    L.ref_sc_detect_closest.argtypes = [C.c_void_p, _pd, _pf, _pd]
    L.ref_sc_detect_closest.restype = C.c_int
    L.ref_sc_detect_candidates.argtypes = [C.c_void_p, _pd, _pi, _pf, _pf, C.c_int]
    L.ref_sc_detect_candidates.restype = C.c_int
    L.ref_sc_ring_knn.argtypes = [C.c_void_p, _pf, _pi, _pf, C.c_int]
    L.ref_sc_ring_knn.restype = C.c_int
    return L


def sc_clouds(n_places=14, revisits=6, seed=7):
    """Sensor-frame clouds (what KeyFrame::computeDescriptor and globalSearch hand to makeScancontext)."""
    from lsdreg import synth
    rng = np.random.default_rng(seed)
    places = []
    for k in range(n_places):
        bi, bj = k % 3, k // 3
        t = synth.block_center(bi, bj) + np.append(rng.uniform(-25, 25, 2), 0.0)
        places.append((bi, bj, t, rng.uniform(-np.pi, np.pi)))
    poses = list(places)
    for k in range(revisits):      # the same place again: another yaw, a metre or two off
        bi, bj, t, yaw = places[int(rng.integers(0, n_places))]
        poses.append((bi, bj, t + np.append(rng.uniform(-1.5, 1.5, 2), 0.0), yaw + rng.uniform(-np.pi, np.pi)))
    clouds = []
    for k, (bi, bj, t, yaw) in enumerate(poses):
        R = synth.rot_from_rpy(0.0, 0.0, yaw)
        clouds.append(np.ascontiguousarray(synth.scan64(100 + k, 180, R, t, bi, bj)))   # scan64 returns sensor-frame points
    return clouds, n_places


def ref_make(L, cloud, dx=0.0, dy=0.0):
    sc, rk, sk = np.zeros(1200), np.zeros(20), np.zeros(60)
    L.ref_sc_make(_p(cloud, _pf), cloud.shape[0], dx, dy, _p(sc, _pd), _p(rk, _pd), _p(sk, _pd))
    return sc.reshape(60, 20), rk, sk


@pytest.fixture(scope="module")
def world(ref):
    clouds, n_places = sc_clouds()
    descs = [S.make(c) for c in clouds]
    return dict(clouds=clouds, descs=descs, n_places=n_places)


def test_descriptor_and_keys_bit_exact(ref, world):
    rng = np.random.default_rng(1)
    for k, c in enumerate(world["clouds"]):
        for dx, dy in ((0.0, 0.0),) + ((S.SEARCH_TRANS[1 + k % 8]),):
            d_ref, rk_ref, sk_ref = ref_make(ref, c, dx, dy)
            d = S.make(c, dx, dy)
            np.testing.assert_array_equal(d, d_ref)
            np.testing.assert_array_equal(S.ringkey(d), rk_ref)
            np.testing.assert_array_equal(S.sectorkey(d), sk_ref)
    # degenerate inputs: empty cloud, points on the axes / at the origin / beyond 80 m / below NO_POINT
    edge = np.array([[0, 0, 1, 0], [5, 0, 1, 0], [0, 5, 2, 0], [-5, 0, 3, 0], [0, -5, 4, 0], [80.0, 0, 1, 0], [56.6, 56.6, 9, 0],
                     [79.9999, 0.01, 2, 0], [3, 3, -2000, 0], [-1e-30, 1e-30, 0.25, 0], [1e-20, -1e-20, 0.5, 0]], np.float32)
    for c in (np.zeros((0, 4), np.float32), edge):
        d_ref, rk_ref, sk_ref = ref_make(ref, c)
        np.testing.assert_array_equal(S.make(c), d_ref)


def test_pairwise_distance_bit_exact(ref, world):
    descs = world["descs"]
    n = len(descs)
    rng = np.random.default_rng(2)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, n, (60, 2))] + [(i, i) for i in range(3)]
    zero = np.zeros((60, 20))
    for a, b in pairs:
        dist, shift = np.zeros(1), np.zeros(1, np.int32)
        A, B = np.ascontiguousarray(descs[a]), np.ascontiguousarray(descs[b])
        ref.ref_sc_distance(_p(A, _pd), _p(B, _pd), _p(dist, _pd), _p(shift, _pi))
        d, s = S.distance(A, B)
        assert s == int(shift[0]), (a, b)
        assert d == float(dist[0]), (a, b, d, dist[0])
    # an empty descriptor: no effective column -> NaN, never below the running minimum
    dist, shift = np.zeros(1), np.zeros(1, np.int32)
    ref.ref_sc_distance(_p(zero, _pd), _p(np.ascontiguousarray(descs[0]), _pd), _p(dist, _pd), _p(shift, _pi))
    d, s = S.distance(zero, descs[0])
    assert d == float(dist[0]) == 10000000.0 and s == int(shift[0]) == 0


def test_retrieval_matches_the_reference(ref, world):
    descs, n_places = world["descs"], world["n_places"]
    db_descs = descs[:n_places]
    flat = np.ascontiguousarray(np.stack([d.reshape(-1) for d in db_descs]))
    found = 0
    for thres in (0.2, 0.6):
        h = ref.ref_sc_db_create(_p(flat, _pd), len(db_descs), thres)
        db = S.Database(db_descs, thres)
        for qi in range(len(descs)):
            for (dx, dy) in S.SEARCH_TRANS[:3]:
                q = S.make(world["clouds"][qi], dx, dy)
                qc = np.ascontiguousarray(q)
                # ring-key candidates: same set, same float distances
                key = S.ringkey(q).astype(np.float32)
                ri, rd = np.zeros(10, np.int32), np.zeros(10, np.float32)
                k = ref.ref_sc_ring_knn(h, _p(key, _pf), _p(ri, _pi), _p(rd, _pf), 10)
                oi, od = db.ring_knn(key)
                assert k == len(oi)
                np.testing.assert_array_equal(np.sort(ri[:k]), np.sort(oi))
                np.testing.assert_array_equal(np.sort(rd[:k]), np.sort(od))
                yaw, score = np.zeros(1, np.float32), np.zeros(1)
                lid = ref.ref_sc_detect_closest(h, _p(qc, _pd), _p(yaw, _pf), _p(score, _pd))
                o_id, o_yaw, o_score = db.detect_closest(q)
                assert (lid, float(yaw[0]), float(score[0])) == (o_id, o_yaw, o_score), (qi, dx, dy)
                ci, cy, cd = np.zeros(10, np.int32), np.zeros(10, np.float32), np.zeros(10, np.float32)
                nc = ref.ref_sc_detect_candidates(h, _p(qc, _pd), _p(ci, _pi), _p(cy, _pf), _p(cd, _pf), 10)
                oc = db.detect_candidates(q)
                assert nc == len(oc)
                assert sorted(zip(ci[:nc].tolist(), cy[:nc].tolist(), cd[:nc].tolist())) == sorted(oc)
                found += lid >= 0
        ref.ref_sc_db_destroy(h)
    assert found > 10   # the revisits and the database entries themselves are found
    # an empty database
    h = ref.ref_sc_db_create(_p(flat, _pd), 0, 0.2)
    yaw, score = np.zeros(1, np.float32), np.zeros(1)
    assert ref.ref_sc_detect_closest(h, _p(np.ascontiguousarray(descs[0]), _pd), _p(yaw, _pf), _p(score, _pd)) == -1
    assert S.Database([]).detect_closest(descs[0])[0] == -1
    ref.ref_sc_db_destroy(h)
