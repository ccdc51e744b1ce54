"""Batched k-NN on the bench map, one process, few launches: the thing to put under ncu.

    python tools/knn_batch_probe.py [n_queries] [--shapes 3,2] [--reps 3]
    ncu --set full --clock-control none --import-source on -k regex:brick_knn_kernel -c 2 -o gpurun_out/X python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 1

Builds the 10.1 M-point map of bench.py in a HashVoxelMap with the brick layout on, then times lsd_knn_query_dev
(5-NN, NEARBY18, d2 < 5) in the requested shapes on random-order and voxel-sorted queries, L2 flushed before each launch.
Prints one JSON line per (shape, order)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lsdreg  # noqa: E402
import torch  # noqa: E402


def opt(name, default):
    return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else default


nq = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1 << 21
shapes = [int(x) for x in opt("--shapes", "3,2").split(",")]
reps = int(opt("--reps", "4"))
if "--lib" in sys.argv:      # A/B builds of the library (same C ABI): python tools/knn_batch_probe.py ... --lib path/to/liblsdreg_variant.so
    lsdreg.capi.lib = lsdreg.capi.load_library(os.path.abspath(opt("--lib", "")))
    lsdreg.lib = lsdreg.capi.lib
lsdreg.init(0)
dev = torch.device("cuda", 0)
synth = bench.load_synth()
m = synth.block_map(bench.MAP_SEED, bench.BLOCKS_X, bench.BLOCKS_Y, bench.SPACING)
g = lsdreg.HashVoxelMap(0.5, 25)
g.enable_bricks(19)
t0 = time.perf_counter()
g.insert(m, 0)
torch.cuda.synchronize()
st, bst = g.stats(), g.brick_stats()
rho = st["points"] / st["cells"]
bpq = 56 + 19 * (8 + 16 * rho)
print(json.dumps({"map_points": st["points"], "voxels": st["cells"], "insert_s": time.perf_counter() - t0, "bricks": bst, "bytes_per_query": bpq}))
rng = np.random.default_rng(99)
q = m[rng.integers(0, m.shape[0], nq)].copy()
q[:, :3] += rng.normal(0.0, 0.1, (nq, 3)).astype(np.float32)
cell = np.round(q[:, :3] / 0.5).astype(np.int64)
order = np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))
stream = torch.cuda.ExternalStream(g.stream(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6581.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
keep = {}
for shape in shapes:
    g.set_knn_shape(shape)
    for name, qq in (("random", q), ("sorted", q[order])):
        qd = torch.from_numpy(np.ascontiguousarray(qq)).to(dev)
        idx = torch.empty((nq, 5), dtype=torch.int32, device=dev)
        d2 = torch.empty((nq, 5), dtype=torch.float32, device=dev)
        cnt = torch.empty(nq, dtype=torch.int32, device=dev)
        ts = []
        for r in range(reps):
            flush.fill_(r)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            g.knn_dev(qd, idx, d2, cnt, k=5, max_sq=5.0)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = float(np.median(ts[1:])) if len(ts) > 1 else ts[0]
        keep[(shape, name)] = (idx, d2, cnt)
        print(json.dumps({"shape": shape, "order": name, "us": us, "all_us": ts, "alg_gbs": nq * bpq / us / 1e3, "frac_of_peak": nq * bpq / us / 1e3 / peak,
                          "found5": float((cnt == 5).float().mean().item())}))
if 2 in shapes and 3 in shapes:
    same = all(torch.equal(a, b) for n in ("random", "sorted") for a, b in zip(keep[(3, n)], keep[(2, n)]))
    print(json.dumps({"bricks_identical_to_thread_shape": bool(same)}))
