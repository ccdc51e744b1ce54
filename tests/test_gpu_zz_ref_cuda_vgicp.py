"""GPU: the VGICP path against the COMPILED reference CUDA matcher (fast_gicp::FastVGICPCuda / FastVGICPCudaCore recompiled
for sm_100a: oracle/ref_cuda_vgicp.cu -> oracle/_ref/libref_cuda_vgicp.so) — the method the reference selects where it is
built with USE_VGICP_CUDA (registrations.cpp:43-55).  The product's VGICP follows the CPU FastVGICP (fast_vgicp_impl.hpp,
double precision, every voxel kept); the CUDA variant computes the same model (compute_derivatives.cu:55-93: weight
sqrt(n_pts), (C_B + R C_A R^T)^-1 frozen at the linearisation point, DIRECT1) in fp32 on the same lossy voxel hash as its
NDT (gaussian_voxelmap.cu: up to 1 % of the points may lose their voxel) and with fp32 atomics in arbitrary order.  Hence
the same bars as tests/test_gpu_ref_cuda.py: costs and derivatives to a few per cent, the aligned pose to 5 mm.

STATUS: the comparator was compiled after this round's GPU budget was spent — neither it nor this test has run on a GPU.
Runs in a subprocess, sorts last, NON-STRICT xfail: it reports xpassed / xfailed and cannot turn the validated suite red.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = os.path.exists(os.path.join(_ROOT, "oracle", "_ref", "libref_cuda_vgicp.so"))

_SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import lsdreg
from lsdreg import synth
from oracle.reg import OracleMatcher, RefVgicpCuda

lsdreg.init(0)
m = synth.block_map(1, 1, 1, 0.25)
m[:, :2] -= np.array([60, 40], np.float32)
Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
tgt = np.array([1.0, -2.0, 1.8])
scan = synth.scan64(2, 200, Rgt, tgt + np.array([60, 40, 0]))[::2].copy()
dR, dt = synth.perturb(5, 0.5, 3.0)
guess = np.eye(4); guess[:3, :3] = Rgt @ dR; guess[:3, 3] = tgt + dt
Tgt = np.eye(4); Tgt[:3, :3] = Rgt; Tgt[:3, 3] = tgt
# registrations.cpp:43-55: eps 0.01, rotation epsilon left at LsqRegistration's default 1e-2 (lsq_registration_impl.hpp:24)
g = lsdreg.Matcher("FAST_VGICP_CUDA", transformation_epsilon=0.01, rotation_epsilon_deg=1e-2)
o = OracleMatcher("vgicp", neighbors=1, trans_eps=0.01, rot_eps=1e-2)
r = RefVgicpCuda(1.0, 64, 0.01, 0)
for mm in (g, o, r):
    mm.set_target(m); mm.set_source(scan)
for name, T in (("guess", guess), ("truth", Tgt)):
    er, Hr, br = r.linearize(T)
    eg, Hg, bg, ncg = g.cost(T)
    eo, Ho, bo = o.cost(T)
    # the reference's two variants may scale the objective differently (a constant factor): compare after normalising by the cost
    s = er / eg
    print(name, "cost ref", er, "ours", eg, "oracle", eo, "ratio", s, "corr", ncg)
    assert 0.2 < s < 5.0
    np.testing.assert_allclose(eg * s, er, rtol=3e-2)
    np.testing.assert_allclose(Hg * s, Hr, rtol=3e-2, atol=3e-2 * np.abs(Hr).max())
    np.testing.assert_allclose(bg * s, br, rtol=3e-2, atol=6e-2 * np.abs(br).max())
    np.testing.assert_allclose(eo, eg, rtol=1e-6)
Tr = r.align(guess)
Tg = g.align(guess)
print("align: |dt| ref vs ours", float(np.abs(Tg[:3, 3] - Tr[:3, 3]).max()), "ours vs truth", float(np.abs(Tg[:3, 3] - tgt).max()),
      "ref vs truth", float(np.abs(Tr[:3, 3] - tgt).max()), "converged", g.converged, r.converged)
assert np.abs(Tg[:3, 3] - Tr[:3, 3]).max() < 5e-3 and np.abs(Tg[:3, :3] - Tr[:3, :3]).max() < 1e-3
print("REF_VGICP_OK")
'''


@pytest.mark.skipif(not HAVE, reason="oracle/_ref/libref_cuda_vgicp.so not built (needs /root/reference + nvcc)")
def test_vgicp_matches_compiled_reference_cuda():
    r = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": _ROOT}], cwd=_ROOT, capture_output=True, text=True, timeout=420)
    tail = (r.stdout[-3000:] + "\n" + r.stderr[-3000:])
    assert r.returncode == 0 and "REF_VGICP_OK" in r.stdout, tail
