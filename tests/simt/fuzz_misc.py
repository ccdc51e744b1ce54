"""TEST INFRASTRUCTURE: adversarial inputs for the voxel grid, the key-frame filters and the ScanContext descriptor, run by
tests/test_emu_kernels.py against the SIMT emulator build (and by hand under AddressSanitizer, tests/simt/README.md)."""
import sys

import numpy as np

import lsdreg
from oracle import filters as OF
from oracle import oracle as O
from oracle import scancontext as S

lsdreg.init(0)
rng = np.random.default_rng(2)


def P(a, w=None):
    a = np.asarray(a, np.float32)
    out = np.zeros((a.shape[0], 4), np.float32)
    out[:, :3] = a
    out[:, 3] = rng.uniform(0, 255, a.shape[0]) if w is None else w
    return out


# ---- voxel grid: sizes 0, 1, one leaf, negative coordinates, leaf borders, a far outlier, a grid that overflows int32
vg = lsdreg.VoxelGrid(max_points=60000, log2_max_cells=28)
cases = {
    "empty": np.zeros((0, 3)),
    "single": [[1.2, -3.4, 0.5]],
    "one leaf": rng.uniform(0.01, 0.49, (300, 3)),
    "negative": rng.uniform(-40, -10, (5000, 3)),
    "borders": np.stack(np.meshgrid(np.arange(-8, 9), np.arange(-8, 9), np.arange(-2, 3), indexing="ij"), -1).reshape(-1, 3) * 0.5,
    "outlier": np.concatenate([rng.uniform(-30, 30, (4000, 3)) * [1, 1, 0.1], [[150.0, -120.0, 20.0]]]),
    "dense": rng.normal(0, 2.0, (50000, 3)),
}
for name, pts in cases.items():
    p = P(pts)
    for leaf in (0.5, 0.2, 1.3):
        got = vg.filter(p, leaf)
        want = O.voxelgrid(p, leaf)
        assert got.shape == want.shape, (name, leaf, got.shape, want.shape)
        if got.shape[0]:
            assert np.abs(got - want).max() <= 2e-4 * max(1.0, float(np.abs(want).max())), (name, leaf, float(np.abs(got - want).max()))
    print("ok voxelgrid", name)
    sys.stdout.flush()
big = P([[0, 0, 0], [9000.0, 9000.0, 900.0], [1, 1, 1]])      # 45000 x 45000 x 4500 leaves at 0.2 m > INT32_MAX: PCL returns the input
got = vg.filter(big, 0.2)
assert got.shape[0] == 3 and (got == big).all(), got
# non-finite coordinates: skipped like pcl::VoxelGrid's !is_dense branch; the finite points give what they give alone
fin = P(rng.uniform(-20, 20, (2000, 3)) * [1, 1, 0.2])
for bad in ([[np.inf, 0, 0]], [[np.nan, 1, 1]], [[0, -np.inf, 0], [1, 1, np.nan], [np.inf, np.inf, np.inf]]):
    mixed = np.concatenate([fin[:700], P(bad), fin[700:]])
    got = vg.filter(mixed, 0.5)
    want = O.voxelgrid(fin, 0.5)
    assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4, (bad, got.shape, want.shape)
assert vg.filter(P([[np.nan, np.nan, np.nan], [np.inf, 0, 0]]), 0.5).shape[0] == 0      # nothing finite: empty output
assert vg.filter(fin, 0.5).shape == O.voxelgrid(fin, 0.5).shape                         # and the scratch is clean afterwards
print("ok voxelgrid non-finite")
small = lsdreg.VoxelGrid(max_points=1000, log2_max_cells=10)   # capacity of the handle exceeded: an error, not a wrong answer
try:
    small.filter(P(rng.uniform(-50, 50, (500, 3))), 0.5)
    raise AssertionError("expected LSD_ERR_CAPACITY")
except lsdreg.LsdError as e:
    assert e.status == lsdreg.ERR_CAPACITY
print("ok voxelgrid limits")

# ---- key-frame filters: order preserved, same kept set as the restatement; tiny and empty inputs; everything an outlier
for name, pts in {"empty": np.zeros((0, 3)), "two": [[1, 1, 0], [1.2, 1.1, 0]], "sparse": rng.uniform(-80, 80, (3000, 3)) * [1, 1, 0.05],
                  "clustered": np.concatenate([rng.normal(0, 1.0, (4000, 3)) + c for c in rng.uniform(-40, 40, (6, 3)) * [1, 1, 0.1]])}.items():
    p = P(pts)
    got = lsdreg.keyframe_filter(p, 1.0, 3, 0.0, 50.0)
    want = OF.keyframe_filter(p, 1.0, 3, 0.0, 50.0)
    assert got.shape == want.shape and (got == want).all(), (name, got.shape, want.shape)
    print("ok keyframe_filter", name, got.shape[0], "of", p.shape[0])

# ---- ScanContext descriptor on degenerate clouds (the retrieval is covered by the zz script)
sc = lsdreg.ScanContext(db_capacity=8)
edge = np.array([[0, 0, 1], [5, 0, 1], [0, 5, 2], [-5, 0, 3], [0, -5, 4], [80.0, 0, 1], [56.6, 56.6, 9], [79.9999, 0.01, 2], [3, 3, -2000],
                 [-1e-30, 1e-30, 0.25], [1e-20, -1e-20, 0.5], [np.nan, 1, 1], [1, np.nan, 1], [np.inf, 1, 1], [-np.inf, -np.inf, 5]], np.float32)
for cloud in (P(edge), P(rng.uniform(-100, 100, (20000, 3)) * [1, 1, 0.05]), P(np.zeros((0, 3)))):
    d, rk, sk = sc.make(cloud, S.SEARCH_TRANS)
    for j, (dx, dy) in enumerate(S.SEARCH_TRANS):
        o = S.make(cloud, dx, dy)
        assert (d[j] == o).all() and (rk[j] == S.ringkey(o)).all() and (sk[j] == S.sectorkey(o)).all(), j
print("ok scancontext descriptors")
print("FUZZ_MISC_OK")
