"""GPU: programmatic dependent launch (lsd_lio_set_pdl, csrc/lsd_common.cuh) changes WHEN a kernel's blocks become resident,
never what they compute: a scan stream registered with the attribute on must give the same bits as with it off — state,
covariance, Nearest_Points ids, downsampled scan, map contents — for the default search shape, the fused shape, with stale
rows, and with the double-buffered ingest (a copy-stream event between two PDL launches).

STATUS: written after this round's GPU budget was spent — it has never run on a GPU.  PDL is OFF by default (without the
launch attribute griddepcontrol.wait / .launch_dependents are no-ops), so nothing else depends on it.  Runs in a subprocess,
sorts last, NON-STRICT xfail: it reports xpassed / xfailed and cannot turn the validated suite red.  Round 2 runs it first
(tools/knn_shapes_probe.py times the stream with and without PDL) and removes the marker.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import lsdreg
from lsdreg import synth
from oracle import eskf

lsdreg.init(0)
m = synth.block_map(1, 2, 2, 0.5)
steps = []
for s in range(6):
    Rgt = synth.rot_from_rpy(0.01 * s, -0.02, 0.3 + 0.05 * s)
    tgt = synth.block_center(0, 0) + np.array([1.0 + 0.4 * s, -2.0 + 0.1 * s, 0.0])
    scan = np.ascontiguousarray(synth.scan64(2 + s, 250 + 37 * (s %% 3), Rgt, tgt), np.float32)     # scans of different sizes: n_down changes from scan to scan
    dR, dt = synth.perturb(5 + s)
    prior = eskf.State(); prior.rot = eskf.R_to_quat(Rgt @ dR); prior.pos = tgt + dt
    steps.append((scan, prior.to_vec()))

def stream(shape, pdl, stale, prefetch):
    f = lsdreg.LioFrontend(map_log2_lines=20, async_map_insert=1 if prefetch else 0)
    f.map.insert(m, 0); f.set_next_id(m.shape[0])
    f.set_knn_shape(shape); f.set_stale_rows(stale); f.set_pdl(pdl)
    out = []
    for s, (scan, x0) in enumerate(steps):
        if prefetch and s + 1 < len(steps):
            f.prefetch(steps[s + 1][0])
        x, P, info = f.scan(scan, x0, lsdreg.init_cov())
        mt = f.get_matches()
        out.append((x.copy(), P.copy(), f.get_down().copy(), mt["idx"].copy(), mt["cnt"].copy() if "cnt" in mt else None,
                    info["n_eff"], info["n_down"], info["iterations"]))
    st = f.map.stats()
    f.close()
    return out, st

for shape, stale, prefetch in ((0, 0, 0), (0, 1, 1), (4, 0, 1), (3, 1, 0)):
    a, sa = stream(shape, 0, stale, prefetch)
    b, sb = stream(shape, 1, stale, prefetch)
    assert sa == sb, (sa, sb)
    for s, (u, v) in enumerate(zip(a, b)):
        for j in range(4):
            np.testing.assert_array_equal(u[j].view(np.int64) if u[j].dtype == np.float64 else u[j].view(np.int32),
                                          v[j].view(np.int64) if v[j].dtype == np.float64 else v[j].view(np.int32),
                                          err_msg=f"shape {shape} stale {stale} prefetch {prefetch} scan {s} field {j}")
        if u[4] is not None:
            np.testing.assert_array_equal(u[4], v[4])
        assert u[5:] == v[5:], (s, u[5:], v[5:])
    print("pdl ok: shape", shape, "stale", stale, "prefetch", prefetch, "n_eff", [u[5] for u in a])
print("PDL_OK")
'''


@pytest.mark.xfail(strict=False, reason="never run on a GPU yet (written after the round's GPU budget was spent); see the module docstring")
def test_pdl_stream_is_bit_identical_to_the_plain_launches():
    r = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": _ROOT}], cwd=_ROOT, capture_output=True, text=True, timeout=420)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0 and "PDL_OK" in r.stdout, tail
