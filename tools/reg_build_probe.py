"""Host-side timing of the matcher's set_target / set_source (where does the build time go?)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lsdreg  # noqa: E402
from lsdreg import synth  # noqa: E402

lsdreg.init(0)
m = synth.block_map(41, 1, 1, 0.22)
m = np.ascontiguousarray(m[np.linspace(0, m.shape[0] - 1, 200000).astype(np.int64)])
Rgt = synth.rot_from_rpy(0.0, 0.0, 0.2)
tgt = synth.block_center(0, 0) + np.array([1.0, -2.0, 0.0])
src = synth.scan64(50, 3800, Rgt, tgt)
src = np.ascontiguousarray(src[np.linspace(0, src.shape[0] - 1, 200000).astype(np.int64)])
guess = np.eye(4); guess[:3, :3] = Rgt; guess[:3, 3] = tgt + [0.2, 0.1, 0.0]
for method in sys.argv[1:] or ["FAST_GICP", "FAST_VGICP"]:
    g = lsdreg.Matcher(method)
    for rep in range(3):
        t0 = time.perf_counter(); g.set_target(m); t1 = time.perf_counter(); g.set_source(src); t2 = time.perf_counter()
        g.align(guess); t3 = time.perf_counter()
        print(method, rep, "set_target %.2f ms  set_source %.2f ms  align %.2f ms  iters %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, g.iterations), flush=True)
