"""TEST INFRASTRUCTURE: random announce / register schedules for the double-buffered ingest and the pipelined voxel grid
(lsd_lio_prefetch, lsd_lio_prefetch_dev, lsd_lio_set_pipeline), run by tests/test_emu_kernels.py against the SIMT emulator
build.  The staging logic is host code — two slots, deferred requests, buffers that change hands, announcements that are
never honoured, scans that arrive unannounced, host and device scans mixed, the switch flipped mid-stream — and that is
exactly what an emulator on one thread CAN show: whatever the schedule, every scan must give the bits of a plain stream.
(Stream / event ordering is not emulated: the GPU run of tests/test_gpu_zz_pdl.py covers that.)"""
import numpy as np

import lsdreg
from lsdreg import synth
from oracle import eskf

lsdreg.init(0)
m = synth.block_map(1, 2, 2, 0.5)
N = 9
steps = []
for s in range(N):
    Rgt = synth.rot_from_rpy(0.005 * s, -0.01, 0.2 + 0.04 * s)
    tgt = synth.block_center(0, 0) + np.array([0.5 + 0.5 * s, -1.0 + 0.2 * s, 0.0])
    scan = np.ascontiguousarray(synth.scan64(20 + s, 90 + 11 * (s % 4), Rgt, tgt), np.float32)
    dR, dt = synth.perturb(40 + s)
    prior = eskf.State(); prior.rot = eskf.R_to_quat(Rgt @ dR); prior.pos = tgt + dt
    steps.append((scan, prior.to_vec()))


class Dev:
    """a device-resident scan under the emulator: "device" memory is host memory"""
    def __init__(self, a):
        self.a = a; self.shape = a.shape; self.is_cuda = True
    def data_ptr(self):
        return self.a.ctypes.data


def run(schedule):
    """schedule: list of ("pre", i, dev) / ("scan", i, dev) / ("pipe", flag) / ("pdl", flag)"""
    f = lsdreg.LioFrontend(map_log2_lines=20, async_map_insert=1)
    f.map.insert(m, 0); f.set_next_id(m.shape[0]); f.set_stale_rows(True)
    host = [sc.copy() for sc, _ in steps]
    dev = [Dev(sc.copy()) for sc, _ in steps]
    out = []
    for op in schedule:
        if op[0] == "pipe":
            f.set_pipeline(op[1])
        elif op[0] == "pdl":
            f.set_pdl(op[1])
        elif op[0] == "pre":
            f.prefetch(dev[op[1]] if op[2] else host[op[1]])
        else:
            i = op[1]
            x, P, info = f.scan(dev[i] if op[2] else host[i], steps[i][1], lsdreg.init_cov())
            mt = f.get_matches()
            out.append((x.copy(), P.copy(), f.get_down().copy(), mt["idx"].copy(), mt["cnt"].copy(), info["n_eff"], info["n_down"], info["iterations"]))
    st, ps = f.map.stats(), f.pipeline_stats()
    f.close()
    return out, st, ps


base, sbase, _ = run([("scan", i, False) for i in range(N)])
assert all(b[5] > 100 for b in base), [b[5] for b in base]

rng = np.random.default_rng(11)
adopted_total = 0
for trial in range(14):
    sched = [("pipe", int(rng.random() < 0.8))]
    for i in range(N):
        r = rng.random()
        d_i = bool(rng.integers(0, 2))
        if r < 0.15:
            sched.append(("pipe", int(rng.random() < 0.7)))
        if r > 0.9:
            sched.append(("pdl", int(rng.integers(0, 2))))
        # announcements before scan i: usually the next scan, sometimes two ahead as well, sometimes a scan that is then
        # registered from the OTHER kind of buffer (the announcement is never honoured), sometimes none
        k = rng.random()
        nxt_dev = bool(rng.integers(0, 2))
        if i + 1 < N and k < 0.7:
            sched.append(("pre", i + 1, nxt_dev))
        if i + 2 < N and k < 0.2:
            sched.append(("pre", i + 2, bool(rng.integers(0, 2))))
        if i + 1 < N and 0.7 <= k < 0.8:
            sched.append(("pre", i + 1, nxt_dev)); sched.append(("pre", i + 1, nxt_dev))      # announced twice
        sched.append(("scan", i, d_i))
    got, st, ps = run(sched)
    assert st == sbase, (trial, st, sbase)
    assert ps["adopted"] <= ps["issued"], ps
    adopted_total += ps["adopted"]
    for s, (u, v) in enumerate(zip(base, got)):
        for j in range(5):
            a, b = u[j], v[j]
            np.testing.assert_array_equal(a.view(np.int64) if a.dtype == np.float64 else a.view(np.int32),
                                          b.view(np.int64) if b.dtype == np.float64 else b.view(np.int32),
                                          err_msg=f"trial {trial} scan {s} field {j}: {sched}")
        assert u[5:] == v[5:], (trial, s, u[5:], v[5:], sched)
    print("trial", trial, "ok", ps)
assert adopted_total > 15, adopted_total      # the schedules did exercise the adoption path
print("FUZZ_PIPELINE_OK")
