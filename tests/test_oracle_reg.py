"""Pins the matcher oracle (oracle/lsd_oracle.c + oracle/reg.py) against the COMPILED reference classes
fast_gicp::FastGICP / FastVGICP / LsqRegistration / se3_exp (oracle/ref_reg.cpp -> oracle/_ref/libref_reg.so;
only PCL's containers and its k-d tree are shimmed).  Runs where the prebuilt oracle/_ref exists."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.HAVE_REF_REG, reason="oracle/_ref/libref_reg.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def scene():
    from lsdreg import synth
    m = synth.block_map(1, 1, 1, 0.5)
    m[:, :2] -= np.array([60, 40], np.float32)
    m = m[(np.abs(m[:, 0]) < 30) & (np.abs(m[:, 1]) < 30)].copy()       # ~15 k points: the compiled reference runs in seconds
    Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
    tgt = np.array([1.0, -2.0, 1.8])
    scan = synth.scan64(2, 200, Rgt, tgt + np.array([60, 40, 0]))
    scan = scan[np.linalg.norm(scan[:, :3], axis=1) < 25][::2].copy()
    dR, dt = synth.perturb(5, 0.3, 2.0)
    guess = np.eye(4); guess[:3, :3] = Rgt @ dR; guess[:3, 3] = tgt + dt
    Tgt = np.eye(4); Tgt[:3, :3] = Rgt; Tgt[:3, 3] = tgt
    return dict(tgt=m, src=scan, guess=guess, Tgt=Tgt)


def test_se3_exp_matches_reference():
    from oracle.reg import ref_se3_exp, se3_exp
    rng = np.random.default_rng(0)
    for scale in (1e-7, 1e-3, 0.3, 2.0):
        for _ in range(5):
            a = rng.normal(0, scale, 6)
            np.testing.assert_allclose(se3_exp(a), ref_se3_exp(a), rtol=0, atol=1e-14)


def _port_covs(nrm):
    return np.eye(3)[None] - 0.999 * nrm[:, :, None] * nrm[:, None, :]


def test_gicp_covariances_and_cost_match_reference(scene):
    from oracle.reg import OracleMatcher, RefMatcher
    o = OracleMatcher("gicp", max_corr=2.0, normal_sq=1e4)
    r = RefMatcher("gicp", max_corr=2.0)
    for mm in (o, r):
        mm.set_target(scene["tgt"]); mm.set_source(scene["src"])
    # PLANE-regularised covariances: U diag(1,1,1e-3) V^T == I - 0.999 n n^T.  A handful of neighbourhoods have two
    # (near-)equal smallest eigenvalues; there the direction is arbitrary in any solver: require 99.5 % agreement.
    for which, nrm in ((0, o.src_nrm), (1, o.tgt_nrm)):
        d = np.abs(_port_covs(nrm) - r.covs(which)).reshape(len(nrm), -1).max(axis=1)
        assert (d < 1e-9).mean() > 0.999, (which, (d < 1e-6).mean())
    for T in (scene["guess"], scene["Tgt"]):
        eo, Ho, bo = o.cost(T)
        er, Hr, br = r.linearize(T)
        np.testing.assert_array_equal(o.corr, r.corr())
        assert o.n_corr == r.n_corr > 1000
        np.testing.assert_allclose(eo, er, rtol=1e-10)
        np.testing.assert_allclose(Ho, Hr, rtol=1e-9, atol=1e-9 * np.abs(Hr).max())
        np.testing.assert_allclose(bo, br, rtol=1e-9, atol=1e-9 * np.abs(br).max())
    T2 = scene["Tgt"].copy(); T2[:3, 3] += [0.03, -0.02, 0.01]
    np.testing.assert_allclose(o.cost(T2, update=False, deriv=False)[0], r.compute_error(T2), rtol=1e-10)


def test_sparse_rings_get_the_true_knn():
    """Round 1 bounded the k = 20 neighbour search of the covariances by a 5 m radius and pinned it only on a scan clipped
    to 25 m, where every point has 20 neighbours that close.  On a whole 64-beam scan the far rings do not, and the
    reference (pcl::search::KdTree::nearestKSearch, fast_gicp_impl.hpp:259) takes the 20 nearest wherever they are — the
    2.4 % cost / 6 % H gap the product showed against BOTH reference variants in round 2's first GPU run
    (tests/test_gpu_zz_ref_cuda_vgicp.py).  Default search radius, unclipped scan: the restatement must equal the compiled
    reference again."""
    from lsdreg import synth
    from oracle.reg import OracleMatcher, RefMatcher
    Rgt = synth.rot_from_rpy(0.01, -0.02, 0.3)
    tgt = np.array([1.0, -2.0, 1.8])
    scan = synth.scan64(2, 200, Rgt, tgt + np.array([60, 40, 0]))[::2].copy()
    assert (np.linalg.norm(scan[:, :3], axis=1) > 40).sum() > 200          # far, sparse rings are in
    o = OracleMatcher("gicp")                                               # normal_sq = 25: the product's default
    r = RefMatcher("gicp")
    for mm in (o, r):
        mm.set_target(scan); mm.set_source(scan)
    d = np.abs(_port_covs(o.src_nrm) - r.covs(0)).reshape(len(o.src_nrm), -1).max(axis=1)
    assert (d < 1e-9).mean() > 0.998, ((d < 1e-9).mean(), (d < 1e-6).mean())


def test_vgicp_voxels_and_cost_match_reference(scene):
    from oracle.reg import OracleMatcher, RefMatcher
    for nb in (1, 7):
        o = OracleMatcher("vgicp", resolution=1.0, neighbors=nb, normal_sq=1e4)
        r = RefMatcher("vgicp", resolution=1.0, neighbors=nb)
        for mm in (o, r):
            mm.set_target(scene["tgt"]); mm.set_source(scene["src"])
        # voxel coordinates and voxel statistics
        rng = np.random.default_rng(3)
        pick = scene["tgt"][rng.integers(0, len(scene["tgt"]), 200), :3].astype(np.float64)
        for p in pick:
            c = np.floor(p / 1.0 - 0.5).astype(int)
            np.testing.assert_array_equal(c, r.coord(p))
            mean, cov = np.zeros(3), np.zeros(9)
            n = O.port.orc_vgicp_voxel(o.vg, int(c[0]), int(c[1]), int(c[2]), mean, cov)
            nr, mr, cr = r.voxel(*c)
            assert n == nr > 0
            np.testing.assert_allclose(mean, mr, rtol=0, atol=1e-9)
            np.testing.assert_allclose(cov.reshape(3, 3), cr, rtol=0, atol=1e-9)
        for T in (scene["guess"], scene["Tgt"]):
            eo, Ho, bo = o.cost(T)
            er, Hr, br = r.linearize(T)
            assert o.n_corr == r.n_corr > 1000
            np.testing.assert_allclose(eo, er, rtol=1e-10)
            np.testing.assert_allclose(Ho, Hr, rtol=1e-9, atol=1e-9 * np.abs(Hr).max())
            np.testing.assert_allclose(bo, br, rtol=1e-9, atol=1e-9 * np.abs(br).max())


@pytest.mark.parametrize("kind", ["gicp", "vgicp"])
def test_lm_loop_matches_reference(scene, kind):
    """Whole align(): same LM trajectory -> final pose within the matcher's own float32 output precision."""
    from oracle.reg import OracleMatcher, RefMatcher
    kw = dict(max_iterations=64, trans_eps=0.01, rot_eps=1e-2) if kind == "gicp" else dict(max_iterations=64, trans_eps=0.1, rot_eps=0.1)
    o = OracleMatcher(kind, resolution=1.0, neighbors=1, max_corr=2.0, normal_sq=1e4, **kw)
    r = RefMatcher(kind, resolution=1.0, neighbors=1, max_corr=2.0, **kw)
    for mm in (o, r):
        mm.set_target(scene["tgt"]); mm.set_source(scene["src"])
    To = o.align(scene["guess"])
    Tr = r.align(scene["guess"])
    assert o.converged == r.converged
    assert np.abs(To[:3, 3] - Tr[:3, 3]).max() < 1e-5, (To[:3, 3], Tr[:3, 3])   # the reference returns a float32 matrix
    assert np.abs(To[:3, :3] - Tr[:3, :3]).max() < 1e-6
    assert np.abs(Tr[:3, 3] - scene["Tgt"][:3, 3]).max() < 0.1
    if kind == "gicp":
        np.testing.assert_allclose(o.fitness(Tr, 25.0), r.fitness(25.0), rtol=1e-4)


def test_fitness_score_matches_the_compiled_calc_fitness_score(scene):
    """Row a16: getFitnessScore itself is PCL's (external), but its in-tree twin — InformationMatrixCalculator::
    calc_fitness_score, the function that weighs every loop edge (information_matrix_calculator.cpp:71-102) — compiles
    unmodified against the k-d tree shim.  The restated score (what the GPU path is held to) must equal it: same float
    transform of the source, same exact nearest neighbour, same `d2 <= max_range` test on the UNSQUARED range, same mean."""
    from oracle.reg import OracleMatcher, ref_calc_fitness_score
    o = OracleMatcher("gicp")
    o.set_target(scene["tgt"]); o.set_source(scene["src"])
    for T in (scene["Tgt"], scene["guess"]):
        for max_range in (25.0, 1.5, 0.05):
            want = ref_calc_fitness_score(scene["tgt"], scene["src"], T, max_range)
            got = o.fitness(T, max_range)
            assert 0 < want < 1e300
            np.testing.assert_allclose(got, want, rtol=1e-12)
    far = np.eye(4); far[:3, 3] = [500.0, 0, 0]
    assert ref_calc_fitness_score(scene["tgt"], scene["src"], far, 1.5) > 1e300 and o.fitness(far, 1.5) > 1e300      # nothing within range: DBL_MAX
