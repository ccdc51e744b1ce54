"""How far do the REFERENCE's own poses move between builds of the reference?  (test infrastructure; needs /root/reference)

bench.py's `pose_parity` compares the product with laserMapping.cpp compiled unmodified (oracle/_ref/libref_fastlio.so: g++ -O3,
x86-64 baseline = SSE2 packets, no FMA).  esti_plane solves a 5x3 system by Eigen's ColPivHouseholderQR in fp32 on world
coordinates, and Eigen's reductions follow the packet width and FMA availability of the target (Core/Redux.h, the product
kernels): the same sources built for another target give other plane coefficients, hence other poses.  This script compiles the
same recipe (oracle/Makefile, libref_fastlio.so) with other flags into oracle/_ref/variants/, registers bench.py's own steps
(same map, same scans, same priors) with each build in a child process, and reports the largest pose difference of every
variant against the shipped build — the reference's own reproducibility floor, to read bench.py's pose_parity against.

    python tools/ref_build_sensitivity.py [--steps 20 --warmup 5] > profiles/r02_ref_build_sensitivity.json
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
FLIO = REF + "/slam/mapping/fastlio"
EIGEN = REF + "/slam/thirdparty/fast_gicp/thirdparty/Eigen"
VARIANTS = {          # name -> flags replacing "-O3 -ffp-contract=off" of the shipped recipe
    "shipped (-O3 -ffp-contract=off, x86-64 baseline: SSE2, no FMA)": None,
    "-O2 (same target)": "-O2 -ffp-contract=off",
    "-O3 -mavx2 (8-wide packets, no FMA)": "-O3 -mavx2 -ffp-contract=off",
    "-O3 -mavx2 -mfma (g++'s default contraction, as on an FMA-baseline target such as aarch64)": "-O3 -mavx2 -mfma",
    "-O3 -march=native": "-O3 -march=native",
}


def build(name, flags):
    if flags is None:
        return os.path.join(ROOT, "oracle", "_ref", "libref_fastlio.so")
    out_dir = os.path.join(ROOT, "oracle", "_ref", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libref_fastlio_%d.so" % (abs(hash(flags)) % 100000))
    o = os.path.join(ROOT, "oracle")
    cmd = (f"g++ -std=c++17 {flags} -fPIC -fopenmp -w -DNDEBUG -DMP_EN -DMP_PROC_NUM=ref_mp_threads -shared -o {out} {o}/ref_fastlio.cpp "
           f"{FLIO}/src/preprocess.cpp {FLIO}/include/ikd-Tree/ikd_Tree.cpp -I{o}/ref_shim_fastlio -I{o}/ref_shim_ikfom -I{FLIO}/src "
           f"-I{FLIO}/include -I{EIGEN}")
    subprocess.run(cmd, shell=True, check=True)
    return out


def child(args):
    import numpy as np
    import bench
    synth = bench.load_synth()
    from oracle import eskf
    from oracle import fastlio as FL
    m = synth.block_map(bench.MAP_SEED, bench.BLOCKS_X, bench.BLOCKS_Y, bench.SPACING)
    ref = FL.RefFastLioBench(capacity=1 << 30, threads=8)
    ref.add_map_points(m)
    poses = []
    for s in range(args.warmup + args.steps):
        scan, Rgt, tgt, Rp, tp = bench.make_step(s)
        x = eskf.State(); x.rot = eskf.R_to_quat(Rp); x.pos = tp.copy()
        x, P, n_down = ref.process_scan(scan, x, eskf.init_P())
        poses.append(x.to_vec()[:7].tolist())
    print(json.dumps(poses))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    import numpy as np
    import bench
    poses = {}
    for name, flags in VARIANTS.items():
        lib = build(name, flags)
        env = dict(os.environ, LSD_REF_FASTLIO_LIB=lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--steps", str(args.steps), "--warmup", str(args.warmup)],
                           env=env, capture_output=True, text=True, cwd=ROOT)
        if r.returncode != 0:
            poses[name] = None
            print(name, "FAILED", r.stderr[-500:], file=sys.stderr)
            continue
        poses[name] = np.array(json.loads(r.stdout.strip().splitlines()[-1]))
    base_name = next(iter(VARIANTS))
    base = poses[base_name]
    out = {"what": __doc__.split("\n\n")[0], "steps_compared": int(base.shape[0]), "bar": "1e-4 m / 1e-5 rad (BASELINE.json north_star)",
           "against": base_name, "variants": {}}
    for name, p in poses.items():
        if name == base_name or p is None:
            continue
        dm, dr = zip(*[bench.pose_delta(a, b) for a, b in zip(base, p)])
        out["variants"][name] = {"max_m": float(max(dm)), "max_rad": float(max(dr)), "median_m": float(np.median(dm)), "median_rad": float(np.median(dr)),
                                 "steps_over_bar": int(sum(1 for a, b in zip(dm, dr) if a > 1e-4 or b > 1e-5))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
