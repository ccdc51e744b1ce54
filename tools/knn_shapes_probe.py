"""A/B of the three k-NN shapes (lsd_knn_set_shape: 1 warp/query, 2 thread/query, 3 flat) on the bench map.
Prints one JSON line per shape (CUDA events on the library stream, L2 flushed before every launch, random and
voxel-sorted query order) and checks that the three shapes return identical bits.  Also times one LIO scan stream
with either per-scan search shape (lsd_lio_set_knn_shape).  Under ncu: `-k regex:knn_query_(flat|thread)_kernel`.

    python tools/knn_shapes_probe.py [n_queries] [--no-lio] [--no-knn] [--shapes 1,2,3] [--map points.npy]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import lsdreg  # noqa: E402
from lsdreg import synth  # noqa: E402

if os.environ.get("LSDREG_EMU"):   # dry run of this script's control flow against the SIMT emulator build (tests/simt)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "simt"))
    import build_emu  # noqa: E402
    lsdreg.capi.lib = lsdreg.capi.load_library(build_emu.build())

def _opt(name, default=None):
    for i, a in enumerate(sys.argv):
        if a == name and i + 1 < len(sys.argv):
            return sys.argv[i + 1]
    return default


_skip = {i + 1 for i, a in enumerate(sys.argv) if a in ("--map", "--shapes")}
args = [a for i, a in enumerate(sys.argv) if i > 0 and not a.startswith("--") and i not in _skip]
nq = int(args[0]) if args else 1 << 21
SHAPES = [(int(x), {1: "warp", 2: "thread", 3: "flat"}[int(x)]) for x in _opt("--shapes", "1,2,3").split(",")]
lsdreg.init(0)
dev = torch.device("cuda", 0)
m = np.load(_opt("--map")) if _opt("--map") else synth.block_map(bench.MAP_SEED, bench.BLOCKS_X, bench.BLOCKS_Y, bench.SPACING)
hmap = None
if "--no-knn" not in sys.argv:
    hmap = lsdreg.HashVoxelMap(0.5, 25)
    hmap.insert(m, 0)
try:
    with open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")) as fh:
        peak = float(json.load(fh)["hbm_gbs"])
except Exception:
    peak = 6650.0   # bench.py's fallback
ALG_BYTES = 680.0   # 56 + 19 (8 + 16 rho), rho = 1.554 on this map

# identical bits first (a smaller batch, host round trip)
rng = np.random.default_rng(5)
qs = m[rng.integers(0, m.shape[0], 200003)].copy()
qs[:, :3] += rng.normal(0.0, 0.1, (qs.shape[0], 3)).astype(np.float32)
ref = None
if hmap is not None:
    hmap.set_knn_shape(2)
    ref = hmap.knn(qs)
for shape, name in ([] if "--no-knn" in sys.argv else SHAPES):
    if shape == 2:
        continue
    hmap.set_knn_shape(shape)
    r = hmap.knn(qs)
    same = all((a.view(np.int32) == b.view(np.int32)).all() for a, b in zip(ref, r))
    print(json.dumps({"shape": name, "identical_to_thread_shape": bool(same), "queries_compared": int(qs.shape[0])}), flush=True)
    assert same, f"shape {shape} differs from the thread-per-query shape"

for shape, name in ([] if "--no-knn" in sys.argv or os.environ.get("LSDREG_EMU") else SHAPES):   # timing needs the GPU
    hmap.set_knn_shape(shape)
    out = bench.run_knn_batch(torch, hmap, m, dev, nq)
    out["shape"] = name
    for order in ("random", "sorted"):
        gbs = nq * ALG_BYTES / (out[order + "_us"] * 1e-6) / 1e9
        out[order + "_alg_GBps"] = round(gbs, 1)
        if peak:
            out[order + "_frac"] = round(gbs / peak, 4)
    print(json.dumps(out), flush=True)
if hmap is not None:
    hmap.set_knn_shape(0)
    hmap.close()

if "--no-lio" not in sys.argv:
    # one scan stream, device time per scan with either search shape (bench.py's step, host-resident scans); the poses of
    # shape 3 must equal the default's bit for bit, those of shape 4 (another reduction tree) to rounding
    steps = [bench.make_step(s) for s in range(12)]
    base = None
    import time
    # (shape, pdl, prefetch, pipeline): warp per point (default) / flat / flat fused with the plane fit and the reduction; the
    # default and the fused shape again with programmatic dependent launch (lsd_lio_set_pdl: same kernels, launch latency
    # hidden); then with the next scan announced (lsd_lio_prefetch: H2D copy of scan k+1 under scan k) and, on top of that,
    # the pipelined voxel grid (lsd_lio_set_pipeline: scan k+1 is also downsampled under scan k)
    pinned = None
    for shape, pdl, pre, pipe in ((0, 0, 0, 0), (3, 0, 0, 0), (4, 0, 0, 0), (0, 1, 0, 0), (4, 1, 0, 0),
                                  (0, 0, 1, 0), (0, 0, 1, 1), (0, 1, 1, 1), (4, 1, 1, 1)):
        f = lsdreg.LioFrontend(map_log2_lines=25, max_scan_points=131072, max_points=100000, async_map_insert=1)
        f.map.insert(m, 0)
        f.set_next_id(m.shape[0])
        f.set_knn_shape(shape)
        f.set_pdl(pdl)
        f.set_pipeline(pipe)
        if pre and pinned is None:
            pinned = [torch.from_numpy(np.ascontiguousarray(st[0], np.float32)).pin_memory() if torch.cuda.is_available()
                      else np.ascontiguousarray(st[0], np.float32) for st in steps]
        ms, wall, errs, poses, launches = [], [], [], [], []
        if pre:
            f.prefetch(pinned[0])
        for s, (scan, Rgt, tgt, Rp, tp) in enumerate(steps):
            t0 = time.perf_counter()
            if pre and s + 1 < len(steps):
                f.prefetch(pinned[s + 1])
            xs, P, info = f.scan(pinned[s] if pre else scan, lsdreg.make_state(pos=tp, rot_xyzw=bench.quat_from_R(Rp)), lsdreg.init_cov())
            t1 = time.perf_counter()
            poses.append(xs.copy())
            if s >= 3:
                ms.append(info["gpu_ms"]); errs.append(float(np.abs(xs[:3] - tgt).max())); launches.append(info["kernel_launches"])
                wall.append((t1 - t0) * 1e3)
        poses = np.array(poses)
        if base is None:
            base = poses
        dmax = float(np.abs(poses - base).max())
        ok = dmax == 0.0 if shape in (0, 3) else dmax < 1e-9
        row = {"lio_knn_shape": shape, "pdl": pdl, "prefetch": pre, "pipeline_vg": pipe, "gpu_ms_per_scan_median": float(np.median(ms)),
               "wall_ms_per_scan_median": float(np.median(wall)), "max_err_m": max(errs),
               "kernel_launches_per_scan": float(np.mean(launches)), "max_abs_state_diff_vs_default": dmax,
               "agrees_with_default": bool(ok)}
        if pipe:
            row["pipeline_stats"] = f.pipeline_stats()
        print(json.dumps(row), flush=True)
        f.close()
