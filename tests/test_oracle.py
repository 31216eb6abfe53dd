"""CPU tests of the oracle (test infrastructure): committed golden vectors, independent witnesses
(scipy cKDTree, numpy SVD / eigh, brute force) and analytic known-answer tests.
PARITY UNPINNED: nothing here is an output of PCL; see oracle/b2icp_oracle.h."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from icpslam_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_golden_nn_integer_fixture(oracle):
    tgt = synth.integer_cloud(11 + 4096, 4096)
    q = synth.integer_cloud(12 + 4096, 4096, unique=False)
    idx, d2 = oracle.KdTree(tgt).nn(q)
    assert sha(idx) == GOLD["nn_integer_4096"]["idx_sha256"]
    assert sha(d2) == GOLD["nn_integer_4096"]["d2_sha256"]
    assert idx[:16].tolist() == GOLD["nn_integer_4096"]["idx_head"]
    bi, bd = oracle.nn_brute(tgt, q)
    assert np.array_equal(bi, idx) and np.array_equal(bd, d2)


def test_golden_tie_heavy_lattice(oracle):
    tgt = synth.integer_cloud(3, 3000, -8, 8)
    q = synth.integer_cloud(4, 20000, -10, 10, unique=False)
    idx, d2 = oracle.KdTree(tgt).nn(q)
    assert sha(idx) == GOLD["nn_lattice_ties"]["idx_sha256"]
    assert sha(d2) == GOLD["nn_lattice_ties"]["d2_sha256"]
    # canonical tie rule: among equal d2 the smallest index (checked against an independent numpy scan)
    for i in range(0, 200):
        d = ((tgt[:, :3].astype(np.float64) - q[i, :3]) ** 2).sum(1)
        assert idx[i] == int(np.flatnonzero(d == d.min())[0])


def test_kdtree_matches_scipy_on_real_valued_clouds(oracle):
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=256)
    idx, d2 = oracle.KdTree(sw[0]).nn(sw[1])
    dd, ii = cKDTree(sw[0][:, :3].astype(np.float64)).query(sw[1][:, :3].astype(np.float64))
    assert (ii == idx).mean() > 0.9999            # float32 vs float64 distances may flip exact ties
    assert np.abs(dd ** 2 - d2).max() < 1e-4
    k_idx, k_d2 = oracle.KdTree(sw[0]).knn(sw[1][:500], 20)
    dd, ii = cKDTree(sw[0][:, :3].astype(np.float64)).query(sw[1][:500, :3].astype(np.float64), k=20)
    assert (ii == k_idx).mean() > 0.999
    assert (np.diff(k_d2, axis=1) >= 0).all()     # sorted=true


def test_knn_edge_cases(oracle):
    pts = synth.as_xyzw(np.random.default_rng(0).uniform(-1, 1, (10, 3)))
    with pytest.raises(RuntimeError):
        oracle.KdTree(pts).knn(pts, 20)           # k > N: PCL's computeCovariances bails out
    idx, d2 = oracle.KdTree(pts).knn(pts, 10)
    assert (idx[:, 0] == np.arange(10)).all() and (d2[:, 0] == 0).all()
    dup = np.concatenate([pts, pts])
    idx, _ = oracle.KdTree(dup).nn(pts)
    assert (idx == np.arange(10)).all()           # duplicates: smallest index


def test_svd3_and_umeyama_kats(oracle):
    rng = np.random.default_rng(1)
    for _ in range(20):
        A = rng.normal(size=(3, 3))
        U, s, V = oracle.svd3(A)
        assert np.abs(U @ np.diag(s) @ V.T - A).max() < 1e-13
        assert np.abs(s - np.linalg.svd(A)[1]).max() < 1e-13
        assert np.abs(U.T @ U - np.eye(3)).max() < 1e-13 and np.abs(V.T @ V - np.eye(3)).max() < 1e-13
    # the GICP covariances run the same recipe with the pair-skip threshold at 1e-15 (where a rotation stops changing
    # an fp64 column; the device kernel restates exactly this): symmetric PSD inputs, as accurate as the 1e-17 variant
    for _ in range(20):
        B = rng.normal(size=(3, 20))
        Cm = np.cov(B) * rng.uniform(1e-4, 10.0)
        U, s, V = oracle.svd3(Cm, cov=True)
        assert np.abs(U @ np.diag(s) @ V.T - Cm).max() < 1e-13 * max(1.0, s[0])
        assert np.abs(s - np.linalg.svd(Cm)[1]).max() < 1e-13 * max(1.0, s[0])
        U17 = oracle.svd3(Cm)[0]
        assert np.abs(np.abs(U.T @ U17) - np.eye(3)).max() < 1e-6   # same singular vectors up to sign
    A = np.outer([1, 2, 3], [4, 5, 6.0])          # rank 1
    U, s, V = oracle.svd3(A)
    assert np.abs(U @ np.diag(s) @ V.T - A).max() < 1e-12 and abs(abs(np.linalg.det(U)) - 1) < 1e-12
    # exact rigid motion, generic cloud
    P = synth.as_xyzw(rng.uniform(-5, 5, (500, 3)))
    T = synth.random_rigid(rng, 2.0, 1.0)
    Q = synth.as_xyzw(P[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    assert np.abs(oracle.umeyama(P, Q) - T).max() < 1e-5
    # planar z = 0 cloud (rank-2 covariance): rotation about z recovered, no reflection
    P2 = P.copy()
    P2[:, 2] = 0
    Tz = np.eye(4)
    Tz[:3, :3] = synth.rot_xyz(0, 0, 0.3)
    Tz[:3, 3] = (0.5, -0.2, 0)
    Q2 = synth.as_xyzw(P2[:, :3].astype(np.float64) @ Tz[:3, :3].T + Tz[:3, 3])
    Tu = oracle.umeyama(P2, Q2)
    assert np.abs(Tu - Tz).max() < 1e-5 and np.linalg.det(Tu[:3, :3]) > 0.999
    # reflection trap: a mirrored cloud must still give a proper rotation (det(U) det(V) < 0 branch)
    Qm = Q.copy()
    Qm[:, 0] *= -1
    assert np.linalg.det(oracle.umeyama(P, Qm)[:3, :3]) > 0.999


def test_transform_cloud_semantics(oracle):
    rng = np.random.default_rng(2)
    cloud = synth.as_xyzw(rng.uniform(-80, 80, (1000, 3)))
    T = synth.random_rigid(rng, 3.0, 0.7)
    outd = oracle.transform_cloud(cloud, T, True)
    ref = (cloud[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    assert np.abs(outd[:, :3] - ref).max() <= 8e-6 and (outd[:, 3] == 1).all()
    outf = oracle.transform_cloud(cloud, T, False)
    assert np.abs(outf[:, :3] - ref).max() <= 3e-5
    Tf = T.astype(np.float32)                      # ((m0 x + m1 y) + m2 z) + m3 in float32
    x, y, z = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    exp = ((Tf[0, 0] * x + Tf[0, 1] * y) + Tf[0, 2] * z) + Tf[0, 3]
    assert np.array_equal(outf[:, 0], exp.astype(np.float32))


def test_golden_p2p_and_kats(oracle):
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    assert sha(sw[0]) == GOLD["c1_p2p"]["cloud_sha256"]          # the generator itself is pinned
    r = oracle.align(oracle.default_params("odometer"), sw[1], sw[0], record_iter=-1)
    assert r["iterations"] == GOLD["c1_p2p"]["iterations"] and r["n_corr"] == GOLD["c1_p2p"]["n_corr"]
    assert np.abs(r["T"] - np.array(GOLD["c1_p2p"]["T"])).max() < 1e-6
    assert sha(r["corr_idx"]) == GOLD["c1_p2p"]["corr_idx_sha256"]
    _, _, sc = synth.planar_stream(3, 2)
    r = oracle.align(oracle.default_params("odometer"), sc[1], sc[0])
    assert r["iterations"] == GOLD["c3_planar_p2p"]["iterations"]
    assert np.abs(r["T"] - np.array(GOLD["c3_planar_p2p"]["T"])).max() < 1e-6
    # analytic: Q = T P, same order, no noise
    rng = np.random.default_rng(0)
    P = synth.as_xyzw(rng.uniform(-5, 5, (2000, 3)))
    T = synth.random_rigid(rng, 0.05, 0.01)
    Q = oracle.transform_cloud(P, T)
    r = oracle.align(oracle.default_params("mapper"), P, Q)
    assert r["converged"] and np.abs(r["T"] - T).max() < 5e-6
    r = oracle.align(oracle.default_params("mapper"), P, P)
    assert r["iterations"] <= 2 and np.abs(r["T"] - np.eye(4)).max() < 1e-7
    # PCL semantics: fewer than 3 correspondences -> converged_ = false
    far = P.copy()
    far[:, :3] += 100
    r = oracle.align(oracle.default_params("odometer"), far, P)
    assert r["rc"] == -4 and not r["converged"]
    assert oracle.align(oracle.default_params("odometer"), P[:0], P)["rc"] == -2
    fit = oracle.fitness(P, Q, T.astype(np.float32))
    assert fit < 1e-9


def test_gicp_covariances_and_golden(oracle):
    rng = np.random.default_rng(3)
    patch = synth.as_xyzw(np.c_[rng.uniform(-1, 1, (400, 2)), 1e-3 * rng.normal(size=400)])
    C = oracle.covariances(patch)
    w, v = np.linalg.eigh(C[0])
    assert np.allclose(w, [1e-3, 1, 1], atol=1e-9) and abs(abs(v[2, 0]) - 1) < 1e-3   # normal = z
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    C = oracle.covariances(sw[0][:512])
    assert abs(np.trace(C, axis1=1, axis2=2).sum() - GOLD["cov_512"]["trace_sum"]) < 1e-6
    assert np.abs(C[0] - np.array(GOLD["cov_512"]["first"])).max() < 1e-9
    # independent witness: neighbourhood covariance via numpy, same regularisation
    idx, _ = oracle.KdTree(sw[0][:512]).knn(sw[0][:8], 20)
    for i in range(8):
        nb = sw[0][:512][idx[i], :3].astype(np.float64)
        w, v = np.linalg.eigh(np.cov(nb.T, bias=True))
        n = v[:, 0]
        assert np.abs(C[i] - (np.eye(3) - (1 - 1e-3) * np.outer(n, n))).max() < 2e-3
    r = oracle.align(oracle.default_params("odometer", oracle.MODE_GICP_BFGS), sw[1], sw[0])
    assert r["rc"] == 0 and r["iterations"] == GOLD["c1_gicp"]["iterations"]
    assert np.abs(r["T"] - np.array(GOLD["c1_gicp"]["T"])).max() < 1e-6
    # analytic KAT for the whole GICP loop
    P = synth.as_xyzw(rng.uniform(-5, 5, (3000, 3)) * [1, 1, 0.2])
    T = synth.random_rigid(rng, 0.05, 0.01)
    Q = oracle.transform_cloud(P, T)
    r = oracle.align(oracle.default_params("mapper", oracle.MODE_GICP_BFGS), P, Q)
    assert r["rc"] == 0 and np.abs(r["T"] - T).max() < 2e-5
    with pytest.raises(RuntimeError):
        oracle.covariances(P[:10])               # N < k_correspondences


def test_pose_algebra_matches_reference_formulas(oracle):
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(4)
    for _ in range(20):
        Ta, Tb = synth.random_rigid(rng, 5, 3), synth.random_rigid(rng, 5, 3)
        a, b = oracle.pose_from_matrix(Ta), oracle.pose_from_matrix(Tb)
        qa = Rot.from_matrix(Ta[:3, :3]).as_quat()                 # x,y,z,w
        assert min(np.abs(a[3:] - qa[[3, 0, 1, 2]]).max(), np.abs(a[3:] + qa[[3, 0, 1, 2]]).max()) < 1e-12
        c = oracle.pose_compose(a, b)
        Tc = Ta @ Tb
        assert np.abs(c[:3] - Tc[:3, 3]).max() < 1e-12
        assert np.abs(Rot.from_quat(c[[4, 5, 6, 3]]).as_matrix() - Tc[:3, :3]).max() < 1e-12
        i = oracle.pose_compose(a, oracle.pose_inverse(a))
        assert np.abs(i[:3]).max() < 1e-12 and abs(abs(i[3]) - 1) < 1e-12


def test_gicp_is_roundoff_sensitive(oracle):
    """Why GICP parity at 1e-4 m needs bit-identical f/df evaluation: PCL's BFGS ends on a
    round-off test (NoProgress), so evaluating the SAME residual in double instead of PCL's float moves the
    final transform by far more than float epsilon.  Both runs still land on the same basin (< 1 cm)."""
    _, _, sw = synth.sweep_sequence(1, 4, n_beams=64, n_az=64)
    worst = 0.0
    for i in range(1, 4):
        p = oracle.default_params("odometer", oracle.MODE_GICP_BFGS)
        a = oracle.align(p, sw[i], sw[i - 1])
        p.reserved[0] = 1
        b = oracle.align(p, sw[i], sw[i - 1])
        d = np.abs(a["T"][:3, 3] - b["T"][:3, 3]).max()
        assert d < 1e-2
        worst = max(worst, d)
    assert worst > 1e-5


def test_map_insert_matches_numpy_restatement(oracle):
    """OctreeMapper::addPointsToMap (reference src/icpslam/octree_mapper.cpp:63-71): one point per voxel, first
    come wins, insertion order kept — the oracle against an independent numpy restatement, incrementally."""
    _, poses, sw = synth.sweep_sequence(9, 3, n_beams=32, n_az=256)
    res = 0.2
    world = [synth.as_xyzw(s[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]) for s, T in zip(sw, poses)]
    m = np.zeros((0, 4), np.float32)
    for w in world:
        m = np.concatenate([m, oracle.map_insert(m, w, res)])
    allp = np.concatenate(world)
    ref = synth.voxel_dedup_first(allp, res)
    assert np.array_equal(m[:, :3], ref[:, :3]) and np.all(m[:, 3] == 1.0)
    key = np.floor(m[:, :3].astype(np.float64) / res).astype(np.int64)
    assert len(np.unique(key, axis=0)) == len(m)                       # at most one point per voxel
    assert len(oracle.map_insert(m, allp, res)) == 0                   # idempotent
    bad = np.array([[np.nan, 0, 0, 1], [1e3, 1e3, 1e3, 1]], np.float32)
    assert np.array_equal(oracle.map_insert(m, bad, res), bad[1:])     # non-finite points are skipped


def test_golden_voxel_filter_and_map_insert(oracle):
    """Regression pins of the two steps either side of the path (made by tests/golden/make_golden.py)."""
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    v = oracle.voxel_filter(sw[0], 0.2)
    assert len(v) == GOLD["voxel_0p2"]["n"] and sha(v) == GOLD["voxel_0p2"]["sha256"]
    m = oracle.map_insert(None, sw[0], 0.2)
    a = oracle.map_insert(m, sw[1], 0.2)
    assert (len(m), len(a)) == (GOLD["map_insert_0p2"]["n_first"], GOLD["map_insert_0p2"]["n_added"])
    assert sha(np.concatenate([m, a])) == GOLD["map_insert_0p2"]["sha256"]


def _numpy_icp_witness(src, tgt, max_iter, max_corr=1.0, eps=1e-6):
    """An independent restatement of pcl::IterativeClosestPoint in numpy / scipy: float32
    in-place transform chain, cKDTree nearest neighbour, gate d2 <= max^2, Umeyama by numpy's SVD in float64,
    DefaultConvergenceCriteria (iterations, incremental rotation / squared translation, absolute MSE change)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt[:, :3].astype(np.float64))
    cur = src[:, :3].astype(np.float32).copy()
    final = np.eye(4, dtype=np.float32)
    prev_mse, it = np.finfo(np.float64).max, 0
    while True:
        d, j = tree.query(cur.astype(np.float64), k=1)
        d2 = ((cur - tgt[j, :3]) ** 2).astype(np.float32)
        d2 = (d2[:, 0] + d2[:, 1]) + d2[:, 2]
        keep = d2.astype(np.float64) <= max_corr * max_corr
        if keep.sum() < 3:
            return final.astype(np.float64), it, False
        P, Q = cur[keep].astype(np.float64), tgt[j[keep], :3].astype(np.float64)
        pc, qc = P.mean(0), Q.mean(0)
        H = (Q - qc).T @ (P - pc) / len(P)
        U, _, Vt = np.linalg.svd(H)
        S = np.diag([1.0, 1.0, -1.0 if np.linalg.det(U) * np.linalg.det(Vt) < 0 else 1.0])
        Rm = U @ S @ Vt
        Tinc = np.eye(4)
        Tinc[:3, :3], Tinc[:3, 3] = Rm, qc - Rm @ pc
        Tf = Tinc.astype(np.float32)
        x, y, z = cur[:, 0].copy(), cur[:, 1].copy(), cur[:, 2].copy()
        for r in range(3):
            cur[:, r] = ((Tf[r, 0] * x + Tf[r, 1] * y) + Tf[r, 2] * z) + Tf[r, 3]
        final = (Tf @ final).astype(np.float32)
        it += 1
        mse = float(d2[keep].astype(np.float64).mean())
        if it >= max_iter:
            return final.astype(np.float64), it, True
        cos_angle = 0.5 * (float(Tf[0, 0] + Tf[1, 1] + Tf[2, 2]) - 1.0)
        tr2 = float(Tf[0, 3] ** 2 + Tf[1, 3] ** 2 + Tf[2, 3] ** 2)
        if (cos_angle >= 1.0 - eps and tr2 <= eps) or abs(mse - prev_mse) < 1e-12:
            return final.astype(np.float64), it, True
        prev_mse = mse


def test_p2p_loop_matches_an_independent_numpy_restatement(oracle):
    """The oracle's whole point-to-point loop (not only its parts) against the numpy witness above, on config-1 and
    config-3 data: same iteration count, same transform to float32 round-off."""
    _, _, sw = synth.sweep_sequence(1, 3, n_beams=64, n_az=64)
    _, _, sc = synth.planar_stream(3, 3)
    for src, tgt, preset, iters in ((sw[1], sw[0], "odometer", 10), (sw[2], sw[1], "mapper", 30), (sc[1], sc[0], "odometer", 10),
                                    (sc[2], sc[1], "mapper", 30)):
        o = oracle.align(oracle.default_params(preset), src, tgt)
        T, it, conv = _numpy_icp_witness(src, tgt, iters)
        assert o["iterations"] == it and bool(o["converged"]) == conv
        assert np.abs(o["T"][:3, 3] - T[:3, 3]).max() < 2e-5 and np.abs(o["T"][:3, :3] - T[:3, :3]).max() < 2e-6


def test_voxel_filter_matches_a_numpy_restatement(oracle):
    """pcl::VoxelGrid::applyFilter restated with numpy: float32 min corner and inverse leaf,
    voxel index x-fastest, float32 sums in input order (np.add.at is sequential), one centroid per voxel in ascending
    index order — bit-identical to the oracle."""
    _, _, sw = synth.sweep_sequence(3, 1, n_beams=64, n_az=256)
    cloud = sw[0]
    for leaf in (0.2, 0.05, 1.0):
        inv = np.float32(1.0) / np.float32(leaf)
        xyz = cloud[:, :3].astype(np.float32)
        min_b = np.floor(xyz.min(axis=0) * inv).astype(np.int64)
        max_b = np.floor(xyz.max(axis=0) * inv).astype(np.int64)
        div = max_b - min_b + 1
        ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int64)
        lin = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
        uniq, inverse, counts = np.unique(lin, return_inverse=True, return_counts=True)
        sums = np.zeros((len(uniq), 3), np.float32)
        np.add.at(sums, inverse, xyz)                       # sequential float32 accumulation, input order
        ref = np.ones((len(uniq), 4), np.float32)
        ref[:, :3] = sums / counts[:, None].astype(np.float32)
        out = oracle.voxel_filter(cloud, leaf)
        assert out.shape == ref.shape and np.array_equal(out, ref)


# ------------------------------------------------------------------------------------------------
# second witness of the GICP oracle: an independent restatement of the outer loop,
# the cost functor and GSL's vector_bfgs2 / linear_minimize (tests/gicp_witness.py)
# ------------------------------------------------------------------------------------------------
def _witness_case(name):
    if name == "C1":      # BASELINE configs[0]: 4k-pt sweep pair, odometer budget (10 outer iterations)
        _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
        return sw[1], sw[0], "odometer"
    _, _, sc = synth.planar_stream(3, 2)  # BASELINE configs[2]: planar 1080-pt scans (rank-2 neighbourhoods)
    return sc[1], sc[0], "odometer"


@pytest.mark.parametrize("case", ["C1", "C3"])
def test_gicp_oracle_and_independent_witness_evaluate_the_same_sequence(oracle, case):
    """Every cost-functor evaluation of the whole align — the point x asked for, f and |gradient| — is bit-identical
    between oracle/gicp_oracle.cpp and tests/gicp_witness.py (numpy + a transcription of GSL's bfgs2 written from
    SURVEY.md App. A.4, not from the oracle), and so are the iteration count and the final transform.  PCL's line
    search ends on a round-off test, so two restatements that differed anywhere in the recursion (bracketing,
    sectioning, cubic / quadratic interpolation, the BFGS direction update, the wrapper's caches, applyState,
    computeRDerivative, the Mahalanobis algebra) would part ways within a few evaluations."""
    import gicp_witness as W
    src, tgt, preset = _witness_case(case)
    p = oracle.default_params(preset, oracle.MODE_GICP_BFGS)
    o, trace = oracle.align_gicp_traced(p, src, tgt)
    assert o["rc"] == 0 and len(trace) > 50
    w = W.gicp_align(src, tgt, oracle.covariances(src), oracle.covariances(tgt), p.max_iterations)
    assert w["iterations"] == o["iterations"] and bool(w["converged"]) == bool(o["converged"])
    assert w["n_corr"] == o["n_corr"]
    assert w["trace"].shape == trace.shape
    assert np.array_equal(w["trace"], trace, equal_nan=True)          # x[6], f, |g| of every evaluation
    assert np.array_equal(w["T"].astype(np.float64), o["T"])


def test_gicp_gradient_matches_finite_differences(oracle):
    """computeRDerivative + the translation part of df against central differences of f (2e-8
    is achievable for the analytic form; f itself is evaluated on float32-transformed points, which bounds what a
    finite difference can resolve)."""
    import gicp_witness as W
    src, tgt, _ = _witness_case("C1")
    cs, ct = oracle.covariances(src), oracle.covariances(tgt)
    tree = cKDTree(tgt[:, :3].astype(np.float64))
    d, nn = tree.query(src[:, :3].astype(np.float64))
    keep = d * d < 1.0
    M = np.tile(np.eye(3), (len(src), 1, 1))
    M[keep] = W.mahalanobis(np.eye(3), cs[keep], ct[nn[keep]])
    fn = W.Functor(src.astype(np.float32), tgt.astype(np.float32), np.where(keep, nn, -1), M, np.eye(4, dtype=np.float32))
    x = [0.02, -0.01, 0.005, 0.004, -0.003, 0.006]
    _, g = fn.fdf(x)
    fd = W.finite_difference_gradient(fn, x, h=2e-4)
    assert np.abs(g - fd).max() <= 2e-3 * max(1.0, np.abs(g).max()), (g, fd)


def test_gicp_covariances_match_a_numpy_restatement(oracle):
    """computeCovariances (k = 20 neighbours, the point included; cov = E[pp^T] - mean mean^T; SVD; singular values
    -> (1, 1, 1e-3)) against scipy's k-d tree + numpy.  Degenerate neighbourhoods aside (the smallest direction
    is then implementation-defined) the regularised matrices agree to round-off."""
    _, _, sw = synth.sweep_sequence(1, 1, n_beams=64, n_az=64)
    cloud = sw[0]
    C = oracle.covariances(cloud)
    xyz = cloud[:, :3].astype(np.float64)
    _, idx = cKDTree(xyz).query(xyz, k=20)
    nb32 = cloud[:, :3][idx]                           # [n, 20, 3] float32
    nb = nb32.astype(np.float64)
    mean = nb.sum(axis=1) / 20.0
    # PCL: `cov(k, l) += pt.k * pt.l` — the product is a FLOAT product, the accumulation is double
    prod = (nb32[:, :, :, None] * nb32[:, :, None, :]).astype(np.float64)
    cov = prod.sum(axis=1) / 20.0 - np.einsum("ni,nj->nij", mean, mean)
    U, S, _ = np.linalg.svd(cov)
    reg = np.einsum("nik,k,njk->nij", U, np.array([1.0, 1.0, 1e-3]), U)
    gap = (S[:, 1] - S[:, 2]) / np.maximum(S[:, 0], 1e-30)     # well separated smallest direction
    ok = gap > 1e-3
    assert ok.mean() > 0.9
    assert np.abs(C[ok] - reg[ok]).max() < 1e-6


def test_compat_octree_restatement_properties(oracle):
    """oracle/octree_oracle.cpp (pcl::octree as OctreeMapper uses it) against what its rules
    imply: one point per voxel of the lattice anchored at (first point - resolution), first come wins, in scan order
    — checked with a numpy restatement; the box grows by new roots; approxNearestSearch returns the point of an
    occupied leaf, is exact for a query sitting on a map point, and is NOT the nearest neighbour in general."""
    _, _, sw = synth.sweep_sequence(9, 3, n_beams=64, n_az=256)
    res = 0.2
    t = oracle.CompatOctree(res)
    added = [t.add_points(sw[0]), t.add_points(sw[1])]
    m = t.points()
    allp = np.concatenate([sw[0], sw[1]])
    p0 = sw[0][0, :3].astype(np.float64)
    org = p0 - res / 2 - res / 2
    key = np.floor((allp[:, :3].astype(np.float64) - org) / res).astype(np.int64)
    _, first = np.unique(key, axis=0, return_index=True)
    first.sort()
    assert sum(added) == len(first) == t.size() and np.array_equal(allp[first], m)
    assert t.add_points(sw[0]) == 0                                   # every voxel of sweep 0 is taken
    mn0, d0 = t.box()
    far = synth.as_xyzw(np.array([[900.0, 0, 0]]))
    assert t.add_points(far) == 1
    mn1, d1 = t.box()
    k = (mn0 - mn1) / res
    assert d1 > d0 and mn1[0] == mn0[0] and mn1[1] <= mn0[1] and np.abs(k - np.round(k)).max() < 1e-6   # roots added towards +x only
    idx = t.approx_nearest(sw[2], key_rule=1)
    assert idx.min() >= 0 and idx.max() < t.size()
    on_map = t.approx_nearest(m[::97], key_rule=1)
    assert np.array_equal(on_map, np.arange(t.size())[::97])          # a query on a map point finds that point
    exact, _ = oracle.nn_brute(t.points(), sw[2])
    assert 0.3 < (idx == exact).mean() < 0.95
    # the literal-1.8 key rule (last child's key handed down) cannot be what ran in the reference's experiments
    bad = t.approx_nearest(sw[2][:2000], key_rule=0)
    mm = t.points()
    assert np.linalg.norm(mm[bad, :3] - sw[2][:2000, :3], axis=1).mean() > 20 * np.linalg.norm(mm[idx[:2000], :3] - sw[2][:2000, :3], axis=1).mean()
