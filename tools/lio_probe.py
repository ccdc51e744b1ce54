"""A/B of launch-side settings of the LIO scan path on the bench workload, one process, one map generation.

    python tools/lio_probe.py "LSD_REUSE_CLUSTER=0" "LSD_REUSE_CLUSTER=8x256" "LSD_REUSE_CLUSTER=8x512,LSD_PDL=0" ...

Each argument is a comma-separated list of environment assignments applied before a fresh LioFrontend is created (the library
reads them in lsd_lio_create).  Per configuration: the bench's value leg (device-resident scans, next scan announced), 3 warm-up
+ 20 timed steps; prints wall / device ms per step, the per-kernel-group profile and the worst position error vs truth."""
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

lsdreg.init(0)
dev = torch.device("cuda", 0)
synth = bench.load_synth()
m = synth.block_map(bench.MAP_SEED, bench.BLOCKS_X, bench.BLOCKS_Y, bench.SPACING)
W, K = 3, 20
steps = [bench.make_step(s) for s in range(W + K)]
scans = [torch.from_numpy(s[0]).to(dev) for s in steps]
P0 = lsdreg.init_cov()
base_env = dict(os.environ)
ref_poses = None
for cfg in (sys.argv[1:] or [""]):
    os.environ.clear(); os.environ.update(base_env)
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("=", 1)
        os.environ[k] = v
    lio = lsdreg.LioFrontend(map_log2_lines=25, max_scan_points=131072, max_points=100000, async_map_insert=1)
    lio.map.insert(m, 0); lio.set_next_id(m.shape[0])
    lio.prefetch(scans[0])
    poses, infos, t0 = [], [], None
    for s, stp in enumerate(steps):
        if s == W:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        if s + 1 < len(steps):
            lio.prefetch(scans[s + 1])
        x, P, info = lio.scan(scans[s], lsdreg.make_state(pos=stp[4], rot_xyzw=bench.quat_from_R(stp[3])), P0)
        poses.append(x[:7].copy())
        if s >= W:
            info["pos_err"] = float(np.abs(x[:3] - stp[2]).max()); infos.append(info)
            if len(infos) >= 2:
                infos[-2]["gpu_ms"] = info["gpu_ms"]
    last_ms, _ = lio.sync(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    infos[-1]["gpu_ms"] = last_ms
    lio.set_profile(True)
    for stp in steps[:5]:
        lio.scan(torch.from_numpy(stp[0]).to(dev), lsdreg.make_state(pos=stp[4], rot_xyzw=bench.quat_from_R(stp[3])), P0)
    prof = lio.get_profile(); lio.set_profile(False)
    if ref_poses is None:
        ref_poses = poses
    same = all(np.array_equal(a, b) for a, b in zip(poses, ref_poses))
    dmax = max(float(np.abs(a - b).max()) for a, b in zip(poses, ref_poses))
    print(json.dumps({"cfg": cfg, "scans_per_s": K / wall, "wall_ms": 1e3 * wall / K, "device_ms": float(np.mean([i["gpu_ms"] for i in infos])),
                      "launches_per_scan": float(np.mean([i["kernel_launches"] for i in infos])), "pos_err_max": max(i["pos_err"] for i in infos),
                      "poses_identical_to_first_cfg": same, "max_pose_diff_vs_first_cfg": dmax,
                      "kernels_us": {k: round(1e3 * v["ms"] / max(v["count"], 1), 2) for k, v in prof.items()}}), flush=True)
    lio.close()
