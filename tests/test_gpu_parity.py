"""GPU parity tests: the CUDA path, called through the C ABI (include/b2icp.h), against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star):
  * integer / index work (nearest-neighbour indices, correspondence indices): bit-exact;
  * float32 squared distances and transformed points: bit-exact (same op order, no FMA);
  * final transform per scan: 1e-4 m translation, 1e-4 rad rotation.
"""
import math

import numpy as np
import pytest

from icpslam_b200 import synth

pytestmark = pytest.mark.gpu

TOL_T = 1e-4    # metres      (north_star)
TOL_R = 1e-4    # radians     (north_star)


def rot_angle(Ra, Rb):
    """Angle of Ra^T Rb from its skew part (well conditioned near 0, unlike acos of the trace)."""
    R = Ra.T @ Rb
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return math.asin(min(1.0, float(np.linalg.norm(v))))


def assert_transform_close(Ta, Tb):
    assert np.abs(Ta[:3, 3] - Tb[:3, 3]).max() <= TOL_T, (Ta[:3, 3], Tb[:3, 3])
    assert rot_angle(Ta[:3, :3], Tb[:3, :3]) <= TOL_R


@pytest.fixture(scope="module")
def R(b2lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return b2lib


# ------------------------------------------------------------------------------------------------
# K1 + K2: grid build + exact NN, bit-identical indices on integer fixtures
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_t,n_q", [(4096, 4096), (65536, 65536), (500000, 65536)])
def test_nn_integer_fixture_bit_exact(R, oracle, n_t, n_q):
    tgt = synth.integer_cloud(11 + n_t, n_t)
    q = synth.integer_cloud(12 + n_q, n_q, unique=False)
    reg = R.Registration(max_correspondence_distance=1e9)
    reg.setInputTarget(tgt)
    idx, d2 = reg.nearestKSearch1(q)
    oi, od = oracle.KdTree(tgt).nn(q)
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2, od)
    if n_t <= 65536:  # exhaustive scan as the second witness
        bi, bd = oracle.nn_brute(tgt, q[:8192])
        assert np.array_equal(idx[:8192], bi) and np.array_equal(d2[:8192], bd)


def test_nn_tie_heavy_lattice_smallest_index_wins(R, oracle):
    tgt = synth.integer_cloud(3, 3000, -8, 8)
    q = synth.integer_cloud(4, 20000, -10, 10, unique=False)
    reg = R.Registration(max_correspondence_distance=1e9)
    reg.setInputTarget(tgt)
    idx, d2 = reg.nearestKSearch1(q)
    bi, bd = oracle.nn_brute(tgt, q)
    assert np.array_equal(idx, bi) and np.array_equal(d2, bd)
    # duplicates in the target (the mapper's NN cloud has them): still the smallest index
    tgt2 = np.concatenate([tgt, tgt[:500]])
    reg.setInputTarget(tgt2)
    idx, d2 = reg.nearestKSearch1(q)
    bi, bd = oracle.nn_brute(tgt2, q)
    assert np.array_equal(idx, bi) and np.array_equal(d2, bd)


def test_nn_real_valued_sweep_vs_oracle(R, oracle):
    _, _, sw = synth.sweep_sequence(1, 2)
    reg = R.Registration()
    reg.setInputTarget(sw[0])
    idx, d2 = reg.nearestKSearch1(sw[1])
    oi, od = oracle.KdTree(sw[0]).nn(sw[1])
    assert np.array_equal(idx, oi)
    assert np.array_equal(d2, od)


def test_nn_far_queries_fall_back_to_exhaustive_scan(R, oracle):
    rng = np.random.default_rng(5)
    tgt = synth.as_xyzw(rng.uniform(-5, 5, (20000, 3)))
    q = synth.as_xyzw(np.concatenate([rng.uniform(-5, 5, (1000, 3)), rng.uniform(200, 900, (300, 3)),
                                      rng.uniform(-1e4, -5e3, (50, 3))]))
    reg = R.Registration()
    reg.setInputTarget(tgt)
    idx, d2 = reg.nearestKSearch1(q)
    bi, bd = oracle.nn_brute(tgt, q)
    assert np.array_equal(idx, bi) and np.array_equal(d2, bd)


def test_nn_planar_and_tiny_clouds(R, oracle):
    _, _, scans = synth.planar_stream(3, 2)
    reg = R.Registration()
    reg.setInputTarget(scans[0])
    idx, d2 = reg.nearestKSearch1(scans[1])
    bi, bd = oracle.nn_brute(scans[0], scans[1])
    assert np.array_equal(idx, bi) and np.array_equal(d2, bd)
    one = synth.as_xyzw(np.array([[1.0, 2.0, 3.0]]))
    reg.setInputTarget(one)
    idx, d2 = reg.nearestKSearch1(scans[1][:10])
    assert (idx == 0).all()
    idx, d2 = reg.nearestKSearch1(scans[1][:0])
    assert len(idx) == 0


# ------------------------------------------------------------------------------------------------
# K6: transforms, bit-exact float results
# ------------------------------------------------------------------------------------------------
def test_transform_cloud_bit_exact(R, oracle):
    rng = np.random.default_rng(7)
    cloud = synth.as_xyzw(rng.uniform(-80, 80, (10001, 3)))
    T = synth.random_rigid(rng, 3.0, 0.7)
    reg = R.Registration()
    assert np.array_equal(reg.transformPointCloud(cloud, T, double=True), oracle.transform_cloud(cloud, T, True))
    assert np.array_equal(reg.transformPointCloud(cloud, T, double=False), oracle.transform_cloud(cloud, T, False))


# ------------------------------------------------------------------------------------------------
# full ICP loop (point-to-point, north_star pipeline)
# ------------------------------------------------------------------------------------------------
def run_pair(R, oracle, src, tgt, preset, **over):
    reg = R.Registration(preset=preset, **over)
    reg.setInputSource(src)
    reg.setInputTarget(tgt)
    aligned = reg.align(want_aligned=True)
    p = oracle.default_params("odometer" if preset == R.PRESET_ODOMETER else "mapper")
    for k, v in over.items():
        setattr(p, k, v)
    o = oracle.align(p, src, tgt, want_aligned=True, record_iter=-1)
    return reg, aligned, o


def test_config1_4k_scan_pair_10_iterations(R, oracle):
    """BASELINE.json configs[0]: single 4k-pt scan vs 4k-pt previous scan, 10 ICP iterations."""
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    assert sw[0].shape == (4096, 4)
    reg, aligned, o = run_pair(R, oracle, sw[1], sw[0], R.PRESET_ODOMETER)
    assert o["rc"] == 0
    assert reg.iterations == o["iterations"]
    assert reg.hasConverged() == bool(o["converged"])
    assert_transform_close(reg.getFinalTransformation(), o["T"])
    idx, d2 = reg.getCorrespondences()
    assert np.array_equal(idx, o["corr_idx"])          # bit-identical correspondence indices
    keep = idx >= 0
    assert np.array_equal(d2[keep], o["corr_d2"][keep])
    assert reg.result.n_corr_last == o["n_corr"]
    assert abs(reg.result.mse_last - o["mse"]) <= 1e-9
    assert np.abs(aligned - o["aligned"]).max() <= 2e-5
    fit = reg.getFitnessScore()
    ofit = oracle.fitness(sw[1], sw[0], reg.getFinalTransformation().astype(np.float32))
    assert abs(fit - ofit) <= 1e-9 * max(1.0, ofit)


def test_config4_unit_64k_scan_pair_30_iterations(R, oracle):
    """One work item of BASELINE.json configs[3]: consecutive 64k-pt sweeps, 30 iterations."""
    _, _, sw = synth.sweep_sequence(4, 2)
    reg, aligned, o = run_pair(R, oracle, sw[1], sw[0], R.PRESET_MAPPER)
    assert reg.iterations == o["iterations"]
    assert_transform_close(reg.getFinalTransformation(), o["T"])
    idx, _ = reg.getCorrespondences()
    assert np.array_equal(idx, o["corr_idx"])


def test_config2_sweep_vs_accumulated_map_through_the_device_map(R, oracle):
    """BASELINE.json configs[1] at reduced map size, end to end on the device: the map is grown by
    addPointsToMap from 12 sweeps in the map frame (one point per 0.2 m voxel), becomes the target without
    leaving the GPU, and the next sweep is registered with 30 iterations; the oracle does the same steps
    (map_insert restatement, then align) on the host."""
    _, poses, sw = synth.sweep_sequence(2, 14, n_az=256)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.resetMap(0.2)
    ref = np.zeros((0, 4), np.float32)
    for k in range(12):
        T = poses[k]
        w = synth.as_xyzw(sw[k][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
        reg.addPointsToMap(w)
        ref = np.concatenate([ref, oracle.map_insert(ref, w, 0.2)])
    assert reg.mapSize() == len(ref) and len(ref) > 50_000
    # the next sweep, placed with its predecessor's pose (the odometry guess of icpslam.cpp:135)
    T = poses[11]
    q = synth.as_xyzw(sw[12][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    reg.setInputTargetFromMap()
    reg.setInputSource(q)
    reg.align()
    o = oracle.align(oracle.default_params("mapper"), q, ref, record_iter=-1)
    assert reg.iterations == o["iterations"]
    assert_transform_close(reg.getFinalTransformation(), o["T"])
    idx, _ = reg.getCorrespondences()
    assert np.array_equal(idx, o["corr_idx"])


def test_config5_localisation_voxel_downsample_50_iterations(R, oracle):
    """BASELINE.json configs[4] at reduced map size: raw sweep -> voxel filter (leaf 0.2) -> 50 ICP iterations
    against a fixed global map; every stage against the oracle."""
    _, poses, sw = synth.sweep_sequence(5, 8, n_az=512)
    reg = R.Registration(preset=R.PRESET_MAPPER, max_iterations=50)
    reg.resetMap(0.2)
    for k in range(7):
        T = poses[k]
        reg.addPointsToMap(synth.as_xyzw(sw[k][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]))
    gmap = reg.mapCloud()
    T = poses[6]
    raw = synth.as_xyzw(sw[7][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    q = reg.voxelFilterCloud(raw, 0.2)
    assert np.array_equal(q, oracle.voxel_filter(raw, 0.2))
    reg.setInputTargetFromMap()
    reg.setInputSource(q)
    reg.align()
    p = oracle.default_params("mapper")
    p.max_iterations = 50
    o = oracle.align(p, q, gmap, record_iter=-1)
    assert reg.iterations == o["iterations"]
    assert_transform_close(reg.getFinalTransformation(), o["T"])
    idx, _ = reg.getCorrespondences()
    assert np.array_equal(idx, o["corr_idx"])


def test_config5_full_size_5m_point_map_properties(R):
    """BASELINE.json configs[4] at FULL map size (5 M points, 64k-point query), through size-independent
    properties: every query's reported neighbour is at the reported distance, no sampled map point is closer
    (exactness on a random subset), and ICP against the big map undoes a known small rigid motion."""
    rng = np.random.default_rng(505)
    n_map, n_q = 5_000_000, 65_536
    # a 600 m x 400 m terrain: gentle hills + walls, about 20 points / m^2
    xy = rng.uniform([-300, -200], [300, 200], (n_map, 2))
    z = 2.0 * np.sin(xy[:, 0] / 30.0) * np.cos(xy[:, 1] / 40.0) + rng.normal(0, 0.02, n_map)
    wall = rng.random(n_map) < 0.15
    z[wall] = rng.uniform(0, 6, wall.sum())
    xy[wall, 0] = np.round(xy[wall, 0] / 25.0) * 25.0 + rng.normal(0, 0.02, wall.sum())
    gmap = np.concatenate([xy, z[:, None], np.ones((n_map, 1))], axis=1).astype(np.float32)
    sel = np.where((np.abs(gmap[:, 0] - 40) < 45) & (np.abs(gmap[:, 1] + 20) < 45))[0]
    src_idx = rng.choice(sel, n_q, replace=False)
    ang = np.radians(0.4)
    Tt = np.eye(4)
    Tt[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    Tt[:3, 3] = [0.12, -0.07, 0.03]
    c = gmap[src_idx, :3].astype(np.float64).mean(axis=0)
    # q = T^-1 applied about the patch centre, plus sensor noise: ICP must find ~T
    Ti = np.linalg.inv(Tt)
    q = (gmap[src_idx, :3].astype(np.float64) - c) @ Ti[:3, :3].T + Ti[:3, 3] + c + rng.normal(0, 0.01, (n_q, 3))
    q = synth.as_xyzw(q)
    reg = R.Registration(preset=R.PRESET_MAPPER, max_iterations=50)
    reg.setInputTarget(gmap)
    idx, d2 = reg.nearestKSearch1(q)
    assert idx.min() >= 0 and idx.max() < n_map
    d = ((q[:, :3].astype(np.float32) - gmap[idx, :3]) ** 2)
    assert np.allclose(d2, d[:, 0] + d[:, 1] + d[:, 2], rtol=1e-5, atol=1e-9)
    probe = rng.choice(n_q, 256, replace=False)                     # exhaustive check of a subset
    for i in probe[:64]:
        dd = ((gmap[:, :3] - q[i, :3]) ** 2).astype(np.float32)
        full = dd[:, 0] + dd[:, 1] + dd[:, 2]
        assert full.min() >= d2[i] * (1 - 1e-6) and full[idx[i]] <= full.min() * (1 + 1e-6)
    reg.setInputSource(q)
    reg.align()
    T = reg.getFinalTransformation()
    # expected transform maps q back onto the map: x -> R (x - c) + t + c
    Texp = np.eye(4)
    Texp[:3, :3] = Tt[:3, :3]
    Texp[:3, 3] = Tt[:3, 3] + c - Tt[:3, :3] @ c
    assert np.abs(T[:3, :3] - Texp[:3, :3]).max() < 2e-4
    assert np.abs(T[:3, 3] - Texp[:3, 3]).max() < 5e-3
    assert reg.iterations <= 50 and reg.hasConverged()


def test_cached_neighbour_certificate_never_changes_a_correspondence(R, oracle):
    """The sweep skips the search of a query whenever its cached-neighbour certificate holds (csrc/nncache.cuh).
    The result must stay the exact nearest neighbour: compare EVERY iteration's correspondences with the oracle
    (max_iterations = k stops both after k sweeps) on tie-free real data, on a tie-heavy integer lattice, on a
    target with duplicated points (the mapper's NN cloud), with a tight and a wide gate."""
    _, _, sw = synth.sweep_sequence(6, 2, n_beams=64, n_az=256)
    lat_t = synth.integer_cloud(11, 6000, -12, 12)
    lat_s = synth.as_xyzw(synth.integer_cloud(12, 4000, -12, 12, unique=False)[:, :3] * 1.0 + np.float32(0.25))
    dup_t = np.concatenate([sw[0], sw[0][::3], sw[0][::7]])
    cases = [(sw[1], sw[0], 1.0), (sw[1], sw[0], 0.3), (sw[1], sw[0], 25.0), (lat_s, lat_t, 1.5), (sw[1], dup_t, 1.0)]
    for src, tgt, gate in cases:
        for k in (1, 2, 3, 5, 8, 13, 30):
            reg, _, o = run_pair(R, oracle, src, tgt, R.PRESET_MAPPER, max_iterations=k, max_correspondence_distance=gate)
            idx, d2 = reg.getCorrespondences()
            assert reg.iterations == o["iterations"], (gate, k)
            assert np.array_equal(idx, o["corr_idx"]), (gate, k, int((idx != o["corr_idx"]).sum()))
            keep = idx >= 0
            assert np.array_equal(d2[keep], o["corr_d2"][keep])
            assert_transform_close(reg.getFinalTransformation(), o["T"])
            if reg.iterations < k:
                break


def test_config3_planar_1080pt_scan_to_scan(R, oracle):
    """BASELINE.json configs[2]: 2-D planar 1080-pt scans (z = 0: rank-2 covariance, nz = 1 grid)."""
    _, _, scans = synth.planar_stream(3, 6)
    for i in range(1, 6):
        reg, _, o = run_pair(R, oracle, scans[i], scans[i - 1], R.PRESET_ODOMETER)
        assert reg.iterations == o["iterations"]
        assert_transform_close(reg.getFinalTransformation(), o["T"])
        T = reg.getFinalTransformation()
        assert abs(T[2, 3]) < 1e-6 and abs(T[2, 2] - 1) < 1e-6


def test_analytic_kat_exact_rigid_motion(R):
    """Q = T*P, same order, no noise: ICP must return T (SURVEY.md §4 KATs)."""
    rng = np.random.default_rng(0)
    P = synth.as_xyzw(rng.uniform(-5, 5, (5000, 3)))
    T = synth.random_rigid(rng, 0.05, 0.01)
    Q = synth.as_xyzw((P[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]))
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.setInputSource(P)
    reg.setInputTarget(Q)
    reg.align()
    assert reg.hasConverged()
    assert np.abs(reg.getFinalTransformation() - T).max() < 5e-6
    # identity
    reg.setInputTarget(P)
    reg.align()
    assert np.abs(reg.getFinalTransformation() - np.eye(4)).max() < 1e-6
    assert reg.iterations <= 2


def test_guess_is_applied_like_pcl(R, oracle):
    _, _, sw = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
    g = synth.random_rigid(np.random.default_rng(3), 0.1, 0.01).astype(np.float32)
    reg = R.Registration()
    reg.setInputSource(sw[1])
    reg.setInputTarget(sw[0])
    reg.align(guess=g)
    o = oracle.align(oracle.default_params("odometer"), sw[1], sw[0], guess=g)
    assert reg.iterations == o["iterations"]
    assert_transform_close(reg.getFinalTransformation(), o["T"])


def test_promote_source_to_target_is_prev_equals_curr(R, oracle):
    """`*prev_cloud_ = *curr_cloud_` (reference icp_odometer.cpp:209) on device, then the next pair."""
    _, _, sw = synth.sweep_sequence(2, 3, n_beams=64, n_az=128)
    reg = R.Registration()
    reg.setInputTarget(sw[0])
    reg.setInputSource(sw[1])
    reg.align()
    reg.promoteSourceToTarget()
    reg.setInputSource(sw[2])
    reg.align()
    o = oracle.align(oracle.default_params("odometer"), sw[2], sw[1])
    assert reg.iterations == o["iterations"]
    assert_transform_close(reg.getFinalTransformation(), o["T"])
    rc, res = reg.alignBatch([sw[1], sw[2]], [sw[0], None], with_fitness=True)
    assert rc == 0
    assert_transform_close(res[1].matrix(), o["T"])
    assert res[1].fitness < 20  # the reference's acceptance test (icp_odometer.cpp:201)


def test_error_codes(R):
    reg = R.Registration()
    cloud = synth.as_xyzw(np.random.default_rng(1).uniform(-1, 1, (100, 3)))
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -9  # NO_SOURCE
    reg.setInputSource(cloud)
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -8  # NO_TARGET
    with pytest.raises(R.B2icpError) as e:
        reg.setInputTarget(cloud[:0])
    assert e.value.code == -2  # EMPTY_CLOUD
    bad = cloud.copy()
    bad[3, 1] = np.nan
    with pytest.raises(R.B2icpError) as e:
        reg.setInputTarget(bad)
    assert e.value.code == -6  # NONFINITE_INPUT
    reg.setInputTarget(cloud)
    reg.setInputSource(bad)
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -6
    # far-apart clouds: fewer than 3 correspondences inside 1 m -> PCL leaves converged_ = false
    far = cloud.copy()
    far[:, :3] += 50
    reg.setInputSource(far)
    with pytest.raises(R.B2icpError) as e:
        reg.align()
    assert e.value.code == -4
    assert not reg.hasConverged()
    with pytest.raises(R.B2icpError) as e:
        R.Registration().getFitnessScore()
    assert e.value.code == -10


# ------------------------------------------------------------------------------------------------
# batched execution (one sweep launch per iteration for many scans)
# ------------------------------------------------------------------------------------------------
def test_batch_shared_target_equals_single_calls(R, oracle):
    """tgt == NULL: every source against the resident target; results identical to one-by-one calls."""
    _, _, sw = synth.sweep_sequence(5, 5, n_beams=64, n_az=256)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.setInputTarget(sw[0])
    singles = []
    for s in sw[1:]:
        reg.setInputSource(s)
        reg.align(raise_on_fail=False)
        singles.append((reg.getFinalTransformation().copy(), reg.iterations))
    rc, res = reg.alignBatch(sw[1:], None, with_fitness=True)
    assert rc == 0
    for (T, it), r in zip(singles, res):
        assert r.iterations == it
        assert np.array_equal(r.matrix(), T)      # same kernels, same order: bit-identical
        assert r.fitness == r.fitness            # not NaN
    o = oracle.align(oracle.default_params("mapper"), sw[2], sw[0])
    assert_transform_close(res[1].matrix(), o["T"])


def test_ragged_batch_and_tiny_sources(R, oracle):
    """Scans of very different sizes in one batch (5 ... 16384 points: slabs with a partial warp, warps past the
    end, a source too small to register) give what the oracle gives scan by scan."""
    _, _, sw = synth.sweep_sequence(5, 2, n_beams=64, n_az=256)
    tgt, full = sw[0], sw[1]
    sizes = [16384, 1000, 37, 33, 32, 31, 5, 2, 4097]
    srcs = [np.ascontiguousarray(full[:: max(1, len(full) // n)][:n]) for n in sizes]
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.setInputTarget(tgt)
    rc, res = reg.alignBatch(srcs, None, with_fitness=True)
    p = oracle.default_params("mapper")
    for s, r in zip(srcs, res):
        o = oracle.align(p, s, tgt)
        assert r.iterations == o["iterations"] and r.converged == int(o["converged"]), len(s)
        assert r.status_detail == o["rc"], (len(s), r.status_detail, o["rc"])
        if o["rc"] == 0:
            assert_transform_close(r.matrix(), o["T"])
    assert rc == -4 and res[7].status_detail == -4          # 2 points: PCL's "not enough correspondences"


def test_streamed_batches_equal_synchronous_batches(R):
    """b2icp_align_batch_submit / _wait (two slot sets, uploads on a copy stream) return, batch by batch and in
    submission order, exactly what b2icp_align_batch returns."""
    _, _, sw = synth.sweep_sequence(3, 9, n_beams=64, n_az=128)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.setInputTarget(sw[0])
    batches = [sw[1:4], sw[4:6], sw[6:9], sw[2:5]]
    rc, ref = reg.alignBatch([s for b in batches for s in b], None, with_fitness=True)
    assert rc == 0
    got = []
    assert reg.alignBatchSubmit(batches[0], with_fitness=True) == 0
    for k in range(1, len(batches)):
        assert reg.alignBatchSubmit(batches[k], with_fitness=True) == 0      # two in flight
        rc, res = reg.alignBatchWait()                                       # the older one
        assert rc == 0
        got += res
    rc, res = reg.alignBatchWait()
    assert rc == 0
    got += res
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert np.array_equal(a.matrix(), b.matrix()) and a.iterations == b.iterations
        assert a.fitness == b.fitness and a.n_corr_last == b.n_corr_last
    # a synchronous batch of 64 or more scans against the handle's target is streamed internally: same results
    flat = [s for b in batches for s in b]
    rc, many = reg.alignBatch([flat[k % len(flat)] for k in range(75)], None, with_fitness=True)
    assert rc == 0 and len(many) == 75
    for k, r in enumerate(many):
        b = ref[k % len(flat)]
        assert np.array_equal(r.matrix(), b.matrix()) and r.iterations == b.iterations and r.fitness == b.fitness
    for k in range(8):
        assert reg.alignBatchSubmit(batches[k % 4]) == 0                     # eight in flight, each on its own stream
    assert reg.alignBatchSubmit(batches[0]) != 0                             # a ninth is refused
    got = []
    for k in range(8):
        rc, res = reg.alignBatchWait()
        assert rc == 0
        got += res
    for a, b in zip(got, ref + ref):
        assert np.array_equal(a.matrix(), b.matrix()) and a.iterations == b.iterations


def test_batch_consecutive_pairs_longer_than_one_chunk(R, oracle):
    """tgt[i] == NULL: pair i registers against src[i-1]; 70 pairs cross the 64-slot chunk border."""
    _, _, sw = synth.sweep_sequence(6, 6, n_beams=32, n_az=128)
    seq = [sw[i % 6] for i in range(71)]
    reg = R.Registration()
    rc, res = reg.alignBatch(seq[1:], [seq[0]] + [None] * 69)
    assert rc == 0
    p = oracle.default_params("odometer")
    for i in (0, 1, 5, 63, 64, 65, 69):
        o = oracle.align(p, seq[i + 1], seq[i])
        assert res[i].iterations == o["iterations"], i
        assert_transform_close(res[i].matrix(), o["T"])


# ------------------------------------------------------------------------------------------------
# K8: voxel-grid downsample (part of BASELINE configs[4])
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("leaf", [0.2, 0.05, 1.0])
def test_voxel_filter_bit_exact(R, oracle, leaf):
    _, _, sw = synth.sweep_sequence(10, 1)
    reg = R.Registration()
    out = reg.voxelFilterCloud(sw[0], leaf)
    ref = oracle.voxel_filter(sw[0], leaf)
    assert out.shape == ref.shape and len(out) < len(sw[0])
    assert np.array_equal(out, ref)            # same voxels, same order, same float centroids
    # idempotent up to re-binning: filtering the centroids again cannot increase the count
    assert len(reg.voxelFilterCloud(out, leaf)) <= len(out)


def test_voxel_filter_edge_cases(R, oracle):
    reg = R.Registration()
    one = synth.as_xyzw(np.array([[1.0, -2.0, 3.0], [1.01, -2.01, 3.01]]))
    assert np.array_equal(reg.voxelFilterCloud(one, 0.5), oracle.voxel_filter(one, 0.5))
    assert len(reg.voxelFilterCloud(one[:0], 0.5)) == 0
    wide = synth.as_xyzw(np.array([[0.0, 0, 0], [5000.0, 5000, 5000]]))
    out = reg.voxelFilterCloud(wide, 0.001)     # voxel count overflows int: PCL returns the input
    assert np.array_equal(out, wide)
    _, _, sc = synth.planar_stream(3, 1)
    assert np.array_equal(reg.voxelFilterCloud(sc[0], 0.1), oracle.voxel_filter(sc[0], 0.1))


# ------------------------------------------------------------------------------------------------
# K9: the mapper's point map (reference src/icpslam/octree_mapper.cpp:56-90)
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_map_insert_bit_exact_and_in_scan_order(R, oracle):
    _, _, sw = synth.sweep_sequence(9, 4, n_beams=64, n_az=256)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.resetMap(0.2)
    ref = np.zeros((0, 4), np.float32)
    for s in sw:
        added = oracle.map_insert(ref, s, 0.2)
        n = reg.addPointsToMap(s)
        assert n == len(added)
        ref = np.concatenate([ref, added])
        assert reg.mapSize() == len(ref)
    assert np.array_equal(reg.mapCloud(), ref)          # same points, same (scan) order
    assert reg.addPointsToMap(sw[0]) == 0               # every voxel of sweep 0 is taken
    reg.resetMap(0.5)                                   # resetMap: new resolution, empty map, table reused
    assert reg.mapSize() == 0
    assert reg.addPointsToMap(sw[1]) == len(oracle.map_insert(None, sw[1], 0.5))


@pytest.mark.gpu
def test_map_insert_edge_cases(R, oracle):
    reg = R.Registration(preset=R.PRESET_MAPPER)
    with pytest.raises(R.B2icpError):
        reg.addPointsToMap(np.zeros((4, 4), np.float32))          # no resetMap yet
    reg.resetMap(0.2)
    assert reg.addPointsToMap(np.zeros((0, 4), np.float32)) == 0
    dup = np.tile(np.array([[1.0, 2.0, 3.0, 1.0]], np.float32), (1000, 1))
    assert reg.addPointsToMap(dup) == 1                           # collisions: one voxel, first point wins
    bad = np.array([[np.nan, 0, 0, 1], [0.01, 0.01, 0.01, 1], [np.inf, 1, 1, 1], [-0.01, 0.01, 0.01, 1]], np.float32)
    assert reg.addPointsToMap(bad) == 2                           # non-finite points are skipped
    m = reg.mapCloud()
    assert np.array_equal(m[1:], bad[[1, 3]]) and np.array_equal(m[0], dup[0])
    # growth across table resizes keeps every earlier voxel occupied
    rng = np.random.default_rng(5)
    big = np.concatenate([rng.uniform(-50, 50, (200_000, 3)).astype(np.float32), np.ones((200_000, 1), np.float32)], axis=1)
    ref = oracle.map_insert(m, big, 0.2)
    assert reg.addPointsToMap(big) == len(ref)
    assert np.array_equal(reg.mapCloud()[3:], ref)


@pytest.mark.gpu
def test_map_nearest_matches_oracle_and_feeds_icp(R, oracle):
    _, _, sw = synth.sweep_sequence(10, 3, n_beams=64, n_az=256)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.resetMap(0.2)
    reg.addPointsToMap(sw[0])
    reg.addPointsToMap(sw[1])
    m = reg.mapCloud()
    idx, nn = reg.approxNearestNeighbors(sw[2])
    oi, _ = oracle.nn_brute(m, sw[2])
    assert np.array_equal(idx, oi)
    assert np.array_equal(nn, m[oi])                    # nn_cloud: map points in query order, duplicates kept
    # the map as the target, device to device: same registration as uploading the downloaded map
    reg.setInputTargetFromMap()
    reg.setInputSource(sw[2])
    reg.align()
    T_map = reg.getFinalTransformation()
    reg2 = R.Registration(preset=R.PRESET_MAPPER)
    reg2.setInputTarget(m)
    reg2.setInputSource(sw[2])
    reg2.align()
    assert np.array_equal(T_map, reg2.getFinalTransformation())


@pytest.mark.gpu
def test_device_resident_mapper_refine_equals_the_step_by_step_path(R, oracle):
    """b2icp_mapper_register / _grow (scan uploaded once, every intermediate cloud on the device) against the same
    sequence made of the separate calls the reference's refineTransformAndGrowMap makes (octree_mapper.cpp:133-173)."""
    from icpslam_b200 import pose6dof
    _, poses, sw = synth.sweep_sequence(12, 4, n_beams=64, n_az=256)
    fused = R.Registration(preset=R.PRESET_MAPPER)
    steps = R.Registration(preset=R.PRESET_MAPPER)
    icp = R.Registration(preset=R.PRESET_MAPPER)
    fused.resetMap(0.2)
    steps.resetMap(0.2)
    T0 = poses[0].astype(np.float32)
    assert fused.mapperGrow(T0, sw[0]) == steps.addPointsToMap(steps.transformPointCloud(sw[0], T0, double=False))
    for k in (1, 2, 3):
        Tr = poses[k - 1].astype(np.float32)                 # the odometry guess: the previous pose
        Tri = np.linalg.inv(poses[k - 1]).astype(np.float32)
        res = fused.mapperRegister(sw[k], Tr, Tri)
        cloud_in_map = steps.transformPointCloud(sw[k], Tr, double=False)
        _, nn_in_map = steps.approxNearestNeighbors(cloud_in_map)
        nn_cloud = steps.transformPointCloud(nn_in_map, Tri, double=False)
        icp.setInputSource(sw[k])
        icp.setInputTarget(nn_cloud)
        icp.align()
        assert np.array_equal(res.matrix(), icp.getFinalTransformation()) and res.iterations == icp.iterations
        o = oracle.align(oracle.default_params("mapper"), sw[k], nn_cloud)
        assert_transform_close(res.matrix(), o["T"])
        Tf = (poses[k - 1] @ res.matrix()).astype(np.float32)
        assert fused.mapperGrow(Tf) == steps.addPointsToMap(steps.transformPointCloud(sw[k], Tf, double=False))
    assert np.array_equal(fused.mapCloud(), steps.mapCloud())


@pytest.mark.gpu
def test_pcl_compat_octree_map_mode_matches_the_oracle(R, oracle):
    """b2icp_map_reset_octree (the PCL-compatible NN mode): lattice anchored on the first point,
    root box grown as pcl::octree grows it, approxNearestNeighbors = PCL's greedy centre-distance descent.  The map
    contents and nn_cloud equal the oracle's restatement (oracle/octree_oracle.cpp), duplicates included; the exact
    mode stays the default and differs."""
    _, poses, sw = synth.sweep_sequence(11, 4, n_beams=64, n_az=256)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.resetMap(0.2, pcl_octree=True)
    tree = oracle.CompatOctree(0.2)
    clouds = []
    for k in range(3):       # sweeps moved into the map frame: the box has to grow in several directions
        T = poses[k]
        w = synth.as_xyzw(sw[k][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
        clouds.append(w)
        assert reg.addPointsToMap(w) == tree.add_points(w)
    far = synth.as_xyzw(np.array([[-400.0, 30, 2], [500.0, -300, 40], [1.0, 2, -90]]))   # roots added towards - and +
    assert reg.addPointsToMap(far) == tree.add_points(far)
    m = reg.mapCloud()
    assert np.array_equal(m, tree.points())                       # same points, same (scan) order
    q = synth.as_xyzw(sw[3][:, :3].astype(np.float64) @ poses[3][:3, :3].T + poses[3][:3, 3])
    idx, nn = reg.approxNearestNeighbors(q)
    oi = tree.approx_nearest(q, key_rule=1)
    assert np.array_equal(idx, oi)                                # PCL's approxNearestSearch, index for index
    assert np.array_equal(nn, m[oi])                              # nn_cloud: map points in query order, duplicates kept
    assert len(np.unique(oi)) < len(oi)
    exact, _ = oracle.nn_brute(m, q)
    assert 0.3 < (oi == exact).mean() < 0.95                      # a greedy descent, not the nearest neighbour
    # the lattice really is PCL's: the exact-mode map (global lattice) holds a different set of points
    reg2 = R.Registration(preset=R.PRESET_MAPPER)
    reg2.resetMap(0.2)
    for w in clouds:
        reg2.addPointsToMap(w)
    reg2.addPointsToMap(far)
    assert not np.array_equal(reg2.mapCloud(), m)
    # refineTransformAndGrowMap on the compat map: the fused device path equals the step-by-step calls
    Tr = poses[2].astype(np.float32)
    Tri = np.linalg.inv(poses[2]).astype(np.float32)
    res = reg.mapperRegister(sw[3], Tr, Tri)
    cim = reg.transformPointCloud(sw[3], Tr, double=False)
    _, nn_in_map = reg.approxNearestNeighbors(cim)
    nn_cloud = reg.transformPointCloud(nn_in_map, Tri, double=False)
    icp = R.Registration(preset=R.PRESET_MAPPER)
    icp.setInputSource(sw[3])
    icp.setInputTarget(nn_cloud)
    icp.align()
    assert np.array_equal(res.matrix(), icp.getFinalTransformation()) and res.iterations == icp.iterations


@pytest.mark.gpu
@pytest.mark.parametrize("sort", ["0", "1"])
def test_nn_search_sorted_and_unsorted_query_paths(R, oracle, monkeypatch, sort):
    """The stand-alone search counting-sorts large query clouds by target cell before the cooperative scan
    (B2ICP_NN_SORT forces either path): both give the oracle's indices and distances, in the caller's order —
    on a real sweep, with far queries (exhaustive fallback), non-finite queries and a ragged tail."""
    monkeypatch.setenv("B2ICP_NN_SORT", sort)
    _, _, sw = synth.sweep_sequence(3, 2)
    rng = np.random.default_rng(9)
    q = np.concatenate([sw[1][:40001], synth.as_xyzw(rng.uniform(300, 500, (37, 3)))])
    q[123, 0] = np.nan
    q[40000, 2] = np.inf
    reg = R.Registration()
    reg.setInputTarget(sw[0])
    idx, d2 = reg.nearestKSearch1(q)
    fin = np.isfinite(q[:, :3]).all(axis=1)
    oi, od = oracle.KdTree(sw[0]).nn(q[fin])
    assert np.array_equal(idx[fin], oi) and np.array_equal(d2[fin], od)
    assert (idx[~fin] == -1).all()
