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
needs_ref = pytest.mark.skipif(not os.path.exists(_REF), reason="oracle/_ref/libref_keyframe.so not built (needs /root/reference)")

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


@needs_ref
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


@needs_ref
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


@needs_ref
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


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    """The kernels' arithmetic (csrc/sc_math.h) compiled for the host: tests/sc_host_harness.cpp."""
    import subprocess
    out = str(tmp_path_factory.mktemp("sc") / "libsc_host_harness.so")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-o", out,
                           os.path.join(_HERE, "sc_host_harness.cpp")])
    H = C.CDLL(out)
    H.h_sc_make.argtypes = [_pf, C.c_int, C.c_double, C.c_double, _pd, _pd, _pf, _pd, _pd]
    H.h_sc_make.restype = None
    H.h_sc_distance.argtypes = [_pd, _pd, _pd, _pi]
    H.h_sc_distance.restype = None
    H.h_sc_ring_knn.argtypes = [_pf, C.c_int, _pf, _pi, _pf]
    H.h_sc_ring_knn.restype = C.c_int
    H.h_sc_yaw.argtypes = [C.c_int]
    H.h_sc_yaw.restype = C.c_float
    return H


@needs_ref
def test_kernel_arithmetic_on_the_host(ref, harness, world):
    """csrc/sc_math.h (what the CUDA kernels execute) against the compiled reference: descriptor, keys, pair distance and
    shift bit for bit; ring-key candidates as a set."""
    H = harness
    descs = []
    for k, c in enumerate(world["clouds"]):
        dx, dy = S.SEARCH_TRANS[k % 9]
        d_ref, rk_ref, sk_ref = ref_make(ref, c, dx, dy)
        d, rk, rkf, vk, nm = np.zeros(1200), np.zeros(20), np.zeros(20, np.float32), np.zeros(60), np.zeros(60)
        H.h_sc_make(_p(c, _pf), c.shape[0], dx, dy, _p(d, _pd), _p(rk, _pd), _p(rkf, _pf), _p(vk, _pd), _p(nm, _pd))
        np.testing.assert_array_equal(d.reshape(60, 20), d_ref)
        np.testing.assert_array_equal(rk, rk_ref)
        np.testing.assert_array_equal(vk, sk_ref)
        np.testing.assert_array_equal(rkf, rk_ref.astype(np.float32))
        descs.append(d.copy())
    edge = np.array([[0, 0, 1, 0], [5, 0, 1, 0], [0, 5, 2, 0], [-5, 0, 3, 0], [0, -5, 4, 0], [80.0, 0, 1, 0], [56.6, 56.6, 9, 0],
                     [79.9999, 0.01, 2, 0], [3, 3, -2000, 0], [-1e-30, 1e-30, 0.25, 0], [1e-20, -1e-20, 0.5, 0],
                     [np.nan, 1, 1, 0], [1, np.nan, 1, 0], [np.inf, 1, 1, 0]], np.float32)
    for c in (np.zeros((0, 4), np.float32), edge):
        d_ref, _, _ = ref_make(ref, c)
        d, rk, rkf, vk, nm = np.zeros(1200), np.zeros(20), np.zeros(20, np.float32), np.zeros(60), np.zeros(60)
        H.h_sc_make(_p(c, _pf), c.shape[0], 0.0, 0.0, _p(d, _pd), _p(rk, _pd), _p(rkf, _pf), _p(vk, _pd), _p(nm, _pd))
        np.testing.assert_array_equal(d.reshape(60, 20), d_ref)
    rng = np.random.default_rng(4)
    n = len(descs)
    zero = np.zeros(1200)
    for a, b in [(int(x), int(y)) for x, y in rng.integers(0, n, (80, 2))] + [(0, 0), (-1, 0), (0, -1)]:
        A = zero if a < 0 else descs[a]
        B = zero if b < 0 else descs[b]
        dr, sr = np.zeros(1), np.zeros(1, np.int32)
        ref.ref_sc_distance(_p(A, _pd), _p(B, _pd), _p(dr, _pd), _p(sr, _pi))
        dh, sh = np.zeros(1), np.zeros(1, np.int32)
        H.h_sc_distance(_p(A, _pd), _p(B, _pd), _p(dh, _pd), _p(sh, _pi))
        assert dh[0] == dr[0] and sh[0] == sr[0], (a, b, dh, dr, sh, sr)
    # ring-key candidates against the reference's nanoflann tree
    n_db = world["n_places"]
    flat = np.ascontiguousarray(np.stack(descs[:n_db]))
    h = ref.ref_sc_db_create(_p(flat, _pd), n_db, 0.2)
    keys = np.ascontiguousarray(np.stack([S.ringkey(d.reshape(60, 20)).astype(np.float32) for d in descs[:n_db]]))
    for qd in descs:
        q = S.ringkey(qd.reshape(60, 20)).astype(np.float32)
        ri, rd = np.zeros(10, np.int32), np.zeros(10, np.float32)
        k = ref.ref_sc_ring_knn(h, _p(q, _pf), _p(ri, _pi), _p(rd, _pf), 10)
        hi, hd = np.zeros(10, np.int32), np.zeros(10, np.float32)
        kh = H.h_sc_ring_knn(_p(keys, _pf), n_db, _p(q, _pf), _p(hi, _pi), _p(hd, _pf))
        assert k == kh
        np.testing.assert_array_equal(np.sort(ri[:k]), np.sort(hi[:k]))
        np.testing.assert_array_equal(np.sort(rd[:k]), np.sort(hd[:k]))
        assert (np.diff(hd[:k]) >= 0).all()
    ref.ref_sc_db_destroy(h)
    for s in range(60):
        assert H.h_sc_yaw(s) == S._deg2rad(s * S.UNIT_SECTORANGLE)


def test_restatement_matches_the_committed_golden_vectors():
    """tests/golden/scancontext_ref.npz (made by the compiled reference, tests/golden/make_golden_scancontext.py) — the
    check that still runs where /root/reference and oracle/_ref do not exist."""
    g = np.load(os.path.join(_HERE, "golden", "scancontext_ref.npz"))
    K = g["clouds_n"].shape[0]
    for k in range(K):
        c = g["clouds"][k][: g["clouds_n"][k]]
        d = S.make(c, *g["offsets"][k])
        np.testing.assert_array_equal(d.reshape(-1), g["desc"][k])
        np.testing.assert_array_equal(S.ringkey(d), g["ringkey"][k])
        np.testing.assert_array_equal(S.sectorkey(d), g["sectorkey"][k])
    for (a, b), dist, shift in zip(g["pairs"], g["pair_dist"], g["pair_shift"]):
        assert S.distance(g["desc"][a].reshape(60, 20), g["desc"][b].reshape(60, 20)) == (dist, shift)
    db = S.Database(g["desc"][: int(g["db_n"])], 0.2)
    for k in range(K):
        lid, yaw, score = db.detect_closest(g["desc"][k].reshape(60, 20))
        assert (lid, yaw, score) == tuple(g["closest"][k]), k
