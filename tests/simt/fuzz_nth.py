# std::nth_element as the search kernel replays it (lsd_debug_nth_element: warp-cooperative for <= 32 elements, serial else and
# past introselect's depth limit) against the oracle's restatement, under the SIMT emulator.  Same body as
# tests/test_gpu_lio.py::test_device_nth_element_is_the_oracles, so that the CPU suite exercises the warp path's logic too.
import ctypes as C
import numpy as np
from oracle import oracle as O
import test_oracle_golden as T

paths = {1: 0, 2: 0, 3: 0}


def check(d, first, nth, last):
    n = d.shape[0]
    want = np.arange(n, dtype=np.int32)
    O.port.orc_nth_element(C.c_void_p(d.ctypes.data), C.c_void_p(want.ctypes.data), n, first, nth, last)
    got, path = lsdreg.capi.debug_nth_element(d, first, nth, last)
    assert (got == want).all(), (n, first, nth, last, path)
    paths[path] += 1


for d, first, nth, last in T._nth_sequences(np.random.default_rng(21), 1200):
    check(d, first, nth, last)
for d, first, nth, last in T._depth_limit_sequences():
    check(d, first, nth, last)
assert paths[1] > 250 and paths[3] > 250 and paths[2] >= 8, paths
print("paths", paths)
print("FUZZ_NTH_OK")
