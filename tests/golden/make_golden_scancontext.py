"""Generates tests/golden/scancontext_ref.npz from the COMPILED REFERENCE (slam/common/Scancontext/Scancontext.cpp built
unmodified into oracle/_ref/libref_keyframe.so by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden_scancontext.py

Contents (seeded synthetic 64-beam scans, thinned to ~2.4 k points each to keep the fixture small):
  clouds [K,N,4] float32 (+ clouds_n), offsets [K,2]      inputs (an offset from the reference's search_trans per cloud)
  desc [K,1200], ringkey [K,20], sectorkey [K,60]         SCManager::makeScancontext / makeRingkey / makeSectorkey
  pairs [P,2], pair_dist [P], pair_shift [P]              SCManager::distanceBtnScanContext(desc[a], desc[b])
  db_n, closest [K,3] = (loop id, yaw, score)             detectClosestMatch against a database of the first db_n descriptors
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_oracle_scancontext as T  # noqa: E402
from oracle import scancontext as S  # noqa: E402

L = C.CDLL(T._REF)   # the same ctypes view the test fixture builds
L.ref_sc_make.argtypes = [T._pf, C.c_int, C.c_double, C.c_double, T._pd, T._pd, T._pd]; L.ref_sc_make.restype = None
L.ref_sc_distance.argtypes = [T._pd, T._pd, T._pd, T._pi]; L.ref_sc_distance.restype = None
L.ref_sc_db_create.argtypes = [T._pd, C.c_int, C.c_double]; L.ref_sc_db_create.restype = C.c_void_p
L.ref_sc_db_destroy.argtypes = [C.c_void_p]
L.ref_sc_detect_closest.argtypes = [C.c_void_p, T._pd, T._pf, T._pd]; L.ref_sc_detect_closest.restype = C.c_int

clouds, n_places = T.sc_clouds(n_places=8, revisits=4, seed=11)
clouds = [np.ascontiguousarray(c[::4]) for c in clouds]
K = len(clouds)
N = max(c.shape[0] for c in clouds)
arr = np.zeros((K, N, 4), np.float32)
cn = np.zeros(K, np.int32)
off = np.zeros((K, 2))
desc, rk, sk = np.zeros((K, 1200)), np.zeros((K, 20)), np.zeros((K, 60))
for k, c in enumerate(clouds):
    arr[k, : c.shape[0]] = c; cn[k] = c.shape[0]
    off[k] = S.SEARCH_TRANS[k % 9]
    d, r, s = T.ref_make(L, c, off[k, 0], off[k, 1])
    desc[k], rk[k], sk[k] = d.reshape(-1), r, s
rng = np.random.default_rng(5)
pairs = rng.integers(0, K, (40, 2)).astype(np.int32)
pd, ps = np.zeros(40), np.zeros(40, np.int32)
for i, (a, b) in enumerate(pairs):
    dd, ss = np.zeros(1), np.zeros(1, np.int32)
    L.ref_sc_distance(T._p(np.ascontiguousarray(desc[a]), T._pd), T._p(np.ascontiguousarray(desc[b]), T._pd), T._p(dd, T._pd), T._p(ss, T._pi))
    pd[i], ps[i] = dd[0], ss[0]
flat = np.ascontiguousarray(desc[:n_places])
h = L.ref_sc_db_create(T._p(flat, T._pd), n_places, 0.2)
closest = np.zeros((K, 3))
for k in range(K):
    yaw, score = np.zeros(1, np.float32), np.zeros(1)
    lid = L.ref_sc_detect_closest(h, T._p(np.ascontiguousarray(desc[k]), T._pd), T._p(yaw, T._pf), T._p(score, T._pd))
    closest[k] = (lid, yaw[0], score[0])
L.ref_sc_db_destroy(h)
out = os.path.join(ROOT, "tests", "golden", "scancontext_ref.npz")
np.savez_compressed(out, clouds=arr, clouds_n=cn, offsets=off, desc=desc, ringkey=rk, sectorkey=sk, pairs=pairs, pair_dist=pd,
                    pair_shift=ps, db_n=np.int32(n_places), closest=closest)
print("wrote", out, os.path.getsize(out), "bytes;", K, "clouds,", int((closest[:, 0] >= 0).sum()), "matches")
