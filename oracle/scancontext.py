"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's ScanContext (row N4 of SURVEY.md section 8f):
slam/common/Scancontext/Scancontext.cpp — makeScancontext :160-203, ring / sector keys :206-236, distDirectSC :77-99,
fastAlignUsingVkey :102-123, distanceBtnScanContext :126-157, detectClosestMatch :268-331, detectCandidateMatch :333-367.

Pinned (tests/test_oracle_scancontext.py) against the unmodified file compiled into oracle/_ref/libref_keyframe.so.
Descriptors are 20 x 60 doubles in Eigen's storage order (column-major: element (ring, sector) at sector * 20 + ring),
here numpy arrays of shape [60, 20] (sector, ring), so `d.reshape(-1)` is MatrixXd::data().

Reductions follow Eigen 3.3's vectorised redux for an SSE2 build (the reference's x86-64 default: packets of two
doubles, two packet accumulators): four interleaved partial sums s_k over i = k mod 4, result (s0 + s2) + (s1 + s3).
"""
from __future__ import annotations

import numpy as np

N_RING, N_SECTOR = 20, 60
MAX_RADIUS = 80.0
LIDAR_HEIGHT = 0.5
NO_POINT = -1000.0
NUM_CANDIDATES = 10
SEARCH_RATIO = 0.1
SC_DIST_THRES = 0.20
UNIT_SECTORANGLE = 360.0 / N_SECTOR
SEARCH_TRANS = [(0, 0), (-4, 0), (4, 0), (0, -4), (0, 4), (-4, -4), (-4, 4), (4, -4), (4, 4)]  # global_localization.cpp:390-392


def eig_sum(v: np.ndarray) -> np.ndarray:
    """Eigen's redux order along the last axis (length a multiple of 4)."""
    n = v.shape[-1]
    assert n % 4 == 0
    g = v.reshape(v.shape[:-1] + (n // 4, 4))
    s = g[..., 0, :].copy()
    for i in range(1, n // 4):
        s = s + g[..., i, :]
    return (s[..., 0] + s[..., 2]) + (s[..., 1] + s[..., 3])


def xy2theta(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Scancontext.cpp:21-36.  x, y float32; `atan` resolves to ::atan(double); the result returns as float."""
    k = 180.0 / np.pi
    with np.errstate(divide="ignore", invalid="ignore"):
        q1 = k * np.arctan((y / x).astype(np.float64))
        q2 = 180.0 - k * np.arctan((y / (-x)).astype(np.float64))
        q3 = 180.0 + k * np.arctan((y / x).astype(np.float64))
        q4 = 360.0 - k * np.arctan(((-y) / x).astype(np.float64))
    out = np.zeros(x.shape, np.float64)
    xp, yp = x >= 0, y >= 0
    out = np.where(xp & yp, q1, out)
    out = np.where(~xp & yp & ~np.isnan(x), q2, out)
    out = np.where((x < 0) & (y < 0), q3, out)
    out = np.where(xp & (y < 0), q4, out)
    return out.astype(np.float32)


def make(xyzi: np.ndarray, dx: float = 0.0, dy: float = 0.0) -> np.ndarray:
    """makeScancontext -> [60, 20] (sector, ring)."""
    p = np.asarray(xyzi, np.float32)
    x = (p[:, 0].astype(np.float64) + dx).astype(np.float32)
    y = (p[:, 1].astype(np.float64) + dy).astype(np.float32)
    z = (p[:, 2].astype(np.float64) + LIDAR_HEIGHT).astype(np.float32)
    rng = np.sqrt((x * x + y * y).astype(np.float64)).astype(np.float32)
    ang = xy2theta(x, y)
    keep = ~(rng.astype(np.float64) > MAX_RADIUS)
    with np.errstate(invalid="ignore"):
        ring_f = np.ceil((rng.astype(np.float64) / MAX_RADIUS) * N_RING)
        sect_f = np.ceil((ang.astype(np.float64) / 360.0) * N_SECTOR)
    ring_f = np.where(~np.isfinite(ring_f), -2147483648.0, ring_f)   # int(NaN) / int(inf) on x86-64 (cvttsd2si -> INT_MIN)
    sect_f = np.where(~np.isfinite(sect_f), -2147483648.0, sect_f)
    ring = np.maximum(np.minimum(N_RING, ring_f.astype(np.int64)), 1) - 1
    sect = np.maximum(np.minimum(N_SECTOR, sect_f.astype(np.int64)), 1) - 1
    desc = np.full((N_SECTOR, N_RING), NO_POINT, np.float64)
    np.maximum.at(desc, (sect[keep], ring[keep]), z[keep].astype(np.float64))
    desc[desc == NO_POINT] = 0.0
    return desc


def ringkey(desc: np.ndarray) -> np.ndarray:
    """makeRingkeyFromScancontext: row means -> [20] double."""
    return eig_sum(np.ascontiguousarray(desc.T)) / float(N_SECTOR)


def sectorkey(desc: np.ndarray) -> np.ndarray:
    """makeSectorkeyFromScancontext: column means -> [60] double."""
    return eig_sum(desc) / float(N_RING)


def fast_align(vkey1: np.ndarray, vkey2: np.ndarray) -> int:
    j = np.arange(N_SECTOR)
    shifts = np.arange(N_SECTOR)[:, None]
    shifted = vkey2[(j[None, :] - shifts) % N_SECTOR]          # circshift right by `shift`
    diff = vkey1[None, :] - shifted
    norms = np.sqrt(eig_sum(diff * diff))
    return int(np.argmin(norms))                                 # strict <: the first minimum


def dist_direct(sc1: np.ndarray, sc2: np.ndarray, shift: int) -> float:
    """distDirectSC(sc1, circshift(sc2, shift))."""
    b = sc2[(np.arange(N_SECTOR) - shift) % N_SECTOR]
    n1 = np.sqrt(eig_sum(sc1 * sc1))
    n2 = np.sqrt(eig_sum(b * b))
    dot = eig_sum(sc1 * b)
    total, num = 0.0, 0
    for c in range(N_SECTOR):
        if n1[c] == 0 or n2[c] == 0:
            continue
        total = total + dot[c] / (n1[c] * n2[c])
        num += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        return float(1.0 - np.float64(total) / np.float64(num))


def distance(sc1: np.ndarray, sc2: np.ndarray):
    """distanceBtnScanContext -> (dist, shift)."""
    a = fast_align(sectorkey(sc1), sectorkey(sc2))
    radius = int(round(0.5 * SEARCH_RATIO * N_SECTOR))
    space = [a]
    for ii in range(1, radius + 1):
        space.append((a + ii + N_SECTOR) % N_SECTOR)
        space.append((a - ii + N_SECTOR) % N_SECTOR)
    best, arg = 10000000.0, 0
    for s in sorted(space):
        d = dist_direct(sc1, sc2, s)
        if d < best:
            best, arg = d, s
    return best, arg


def ring_d2(keys: np.ndarray, q: np.ndarray) -> np.ndarray:
    """nanoflann L2_Adaptor::evalMetric in float: groups of four, ((d0^2 + d1^2) + d2^2) + d3^2 added to the running sum."""
    keys = np.asarray(keys, np.float32)
    d = (q.astype(np.float32)[None, :] - keys)
    sq = d * d
    res = np.zeros(keys.shape[0], np.float32)
    for g in range(0, N_RING, 4):
        res = res + (((sq[:, g] + sq[:, g + 1]) + sq[:, g + 2]) + sq[:, g + 3])
    return res


def _deg2rad(deg: float) -> float:
    """Scancontext.cpp:15-18: float deg2rad(float degrees) { return degrees * M_PI / 180.0; }"""
    return float(np.float32(float(np.float32(deg)) * np.pi / 180.0))


class Database:
    """buildRingKeyKDTree + detectClosestMatch / detectCandidateMatch (exact k-NN on the float ring keys)."""

    def __init__(self, descs, dist_thres: float = SC_DIST_THRES):
        self.descs = [np.asarray(d, np.float64).reshape(N_SECTOR, N_RING) for d in descs]
        self.keys = np.array([ringkey(d).astype(np.float32) for d in self.descs], np.float32).reshape(-1, N_RING)
        self.thres = dist_thres

    def ring_knn(self, key: np.ndarray):
        k = min(NUM_CANDIDATES, len(self.descs))
        d2 = ring_d2(self.keys, key)
        order = np.lexsort((np.arange(d2.shape[0]), d2))[:k]
        return order, d2[order]

    def candidates(self, sc: np.ndarray):
        """-> list of (db index, dist, shift) for the ring-key candidates, in candidate order."""
        if not self.descs:
            return []
        idx, _ = self.ring_knn(ringkey(sc).astype(np.float32))
        return [(int(i),) + distance(sc, self.descs[int(i)]) for i in idx]

    def detect_closest(self, sc: np.ndarray):
        """-> (loop id or -1, yaw rad, score)."""
        if not self.descs:
            return -1, 0.0, 1.0
        best, align, nn = 10000000.0, 0, 0
        for i, d, s in self.candidates(sc):
            if d < best:
                best, align, nn = d, s, i
        yaw = _deg2rad(align * UNIT_SECTORANGLE)
        return (nn if best < self.thres else -1), yaw, best

    def detect_candidates(self, sc: np.ndarray):
        out = []
        for i, d, s in self.candidates(sc):
            if d < self.thres:
                out.append((i, _deg2rad(s * UNIT_SECTORANGLE), float(np.float32(d))))
        return out
