"""Full-size GPU parity against the CPU oracle on the BASELINE.json configurations themselves (not reduced copies):
configs[1] 64k-pt sweep vs the 500k-pt map (point-to-point and GICP), configs[3] 33 consecutive 64k-pt sweeps through
the replay path, configs[4] 64k-pt sweep vs a 5M-pt map with the voxel filter and 50 iterations.  The oracle's k-d
tree handles each in seconds.  Bars: correspondence indices bit-identical (SHA-256 of the index array), iteration
counts equal, final transform within 1e-4 m / 1e-4 rad (GICP: bit-identical to the oracle's arithmetic definition)."""
import hashlib

import numpy as np
import pytest

from icpslam_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rot_angle(Ra, Rb):
    R = Ra.T @ Rb
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, float(np.linalg.norm(v)))))


def assert_close(Ta, Tb):
    assert np.abs(Ta[:3, 3] - Tb[:3, 3]).max() <= TOL, (Ta[:3, 3], Tb[:3, 3])
    assert rot_angle(Ta[:3, :3], Tb[:3, :3]) <= TOL


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def R(b2lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return b2lib


@pytest.fixture(scope="module")
def config1():
    """BASELINE configs[1] as bench.py builds it: the 500 000-point accumulated map and the next sweeps."""
    m = synth.build_local_map(2, n_points=500_000)
    qs, _ = synth.map_queries(m, 2, 0, 3)
    assert m["map"].shape == (500_000, 4) and qs[0].shape == (65_536, 4)
    return m["map"], qs


def test_config1_full_size_p2p_vs_oracle(R, oracle, config1):
    gmap, qs = config1
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reg.setInputTarget(gmap)
    p = oracle.default_params("mapper")
    for q in qs[:2]:
        reg.setInputSource(q)
        reg.align()
        o = oracle.align(p, q, gmap, record_iter=-1)
        assert reg.iterations == o["iterations"] and reg.hasConverged() == bool(o["converged"])
        assert_close(reg.getFinalTransformation(), o["T"])
        idx, d2 = reg.getCorrespondences()
        assert sha(idx) == sha(o["corr_idx"])                      # 65 536 correspondence indices, bit-identical
        keep = idx >= 0
        assert np.array_equal(d2[keep], o["corr_d2"][keep])
        assert reg.result.n_corr_last == o["n_corr"]
    # the same sweeps as one batch and as a streamed batch: the same bits
    rc, res = reg.alignBatch(qs[:2])
    assert rc == 0
    reg.setInputSource(qs[1])
    reg.align()
    assert np.array_equal(res[1].matrix(), reg.getFinalTransformation()) and res[1].iterations == reg.iterations


def test_config1_full_size_gicp_vs_oracle(R, oracle, config1):
    """The reference's own estimator (GICP, icp_odometer.cpp:188 / octree_mapper.cpp:104) at full size: 64k vs 500k and
    64k vs 64k, final transform bit-identical to the oracle's arithmetic definition, and the batch path equal to it."""
    gmap, qs = config1
    p = oracle.default_params("mapper", oracle.MODE_GICP_BFGS)
    reg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS)
    reg.setInputTarget(gmap)
    reg.setInputSource(qs[0])
    reg.align()
    o = oracle.align(p, qs[0], gmap, record_iter=-1)
    assert reg.iterations == o["iterations"] and reg.result.n_corr_last == o["n_corr"]
    assert np.array_equal(reg.getFinalTransformation(), o["T"])
    idx, _ = reg.getCorrespondences()
    assert sha(idx) == sha(o["corr_idx"])
    rc, res = reg.alignBatch(qs[:3])                                   # three scans advancing in lockstep rounds
    assert rc == 0 and np.array_equal(res[0].matrix(), o["T"]) and res[0].iterations == o["iterations"]
    o2 = oracle.align(p, qs[2], gmap)
    assert np.array_equal(res[2].matrix(), o2["T"]) and res[2].iterations == o2["iterations"]
    # 64k vs 64k (configs[3] unit) in the odometer's budget
    _, _, sw = synth.sweep_sequence(4, 2)
    po = oracle.default_params("odometer", oracle.MODE_GICP_BFGS)
    reg2 = R.Registration(preset=R.PRESET_ODOMETER, mode=R.MODE_GICP_BFGS)
    reg2.setInputTarget(sw[0])
    reg2.setInputSource(sw[1])
    reg2.align()
    o3 = oracle.align(po, sw[1], sw[0])
    assert np.array_equal(reg2.getFinalTransformation(), o3["T"]) and reg2.iterations == o3["iterations"]


def test_config3_33_consecutive_64k_sweeps_through_replay(R, oracle):
    """BASELINE configs[3] at sweep size: 33 recorded 64k-pt sweeps, pair i = sweep i vs sweep i-1, through
    icpslam_b200.replay (b2icp_align_batch in consecutive mode + fitness + the serial fix-up and composition)."""
    from icpslam_b200 import replay
    _, poses, sw = synth.sweep_sequence(4, 33)
    reg = R.Registration(preset=R.PRESET_MAPPER)
    records, traj = replay.replay_pairs(sw, reg)
    assert records.shape == (32, replay.RECORD) and traj.shape == (33, 7)
    assert all(replay.accepted(r) for r in records)
    p = oracle.default_params("mapper")
    for i in (0, 7, 19, 31):
        o = oracle.align(p, sw[i + 1], sw[i], record_iter=-1)
        assert int(records[i, 17]) == o["iterations"], i
        assert_close(records[i, :16].reshape(4, 4), o["T"])
        ofit = oracle.fitness(sw[i + 1], sw[i], records[i, :16].reshape(4, 4).astype(np.float32))
        assert abs(records[i, 20] - ofit) <= 1e-9 * max(1.0, ofit)
    # the composed trajectory follows the ground truth of the synthetic drive (sensor noise 2 cm)
    true_rel = [np.linalg.inv(poses[i]) @ poses[i + 1] for i in range(32)]
    err = max(np.abs(records[i, :16].reshape(4, 4)[:3, 3] - true_rel[i][:3, 3]).max() for i in range(32))
    assert err < 0.3


def test_config4_full_size_5m_map_vs_oracle(R, oracle):
    """BASELINE configs[4] at full size: a raw sweep -> voxel filter (leaf 0.2) -> 50 iterations against a 5 000 000-point
    map, every stage against the oracle (its k-d tree over 5 M points builds in seconds)."""
    rng = np.random.default_rng(505)
    n_map = 5_000_000
    xy = rng.uniform([-300, -200], [300, 200], (n_map, 2))
    z = 2.0 * np.sin(xy[:, 0] / 30.0) * np.cos(xy[:, 1] / 40.0) + rng.normal(0, 0.02, n_map)
    wall = rng.random(n_map) < 0.15
    z[wall] = rng.uniform(0, 6, wall.sum())
    xy[wall, 0] = np.round(xy[wall, 0] / 25.0) * 25.0 + rng.normal(0, 0.02, wall.sum())
    gmap = np.concatenate([xy, z[:, None], np.ones((n_map, 1))], axis=1).astype(np.float32)
    sel = np.where((np.abs(gmap[:, 0] - 40) < 45) & (np.abs(gmap[:, 1] + 20) < 45))[0]
    src_idx = rng.choice(sel, 130_000, replace=False)
    ang = np.radians(0.4)
    Tt = np.eye(4)
    Tt[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    Tt[:3, 3] = [0.12, -0.07, 0.03]
    c = gmap[src_idx, :3].astype(np.float64).mean(axis=0)
    Ti = np.linalg.inv(Tt)
    raw = synth.as_xyzw((gmap[src_idx, :3].astype(np.float64) - c) @ Ti[:3, :3].T + Ti[:3, 3] + c + rng.normal(0, 0.01, (130_000, 3)))
    reg = R.Registration(preset=R.PRESET_MAPPER, max_iterations=50)
    q = reg.voxelFilterCloud(raw, 0.2)
    assert np.array_equal(q, oracle.voxel_filter(raw, 0.2))              # K8 at full size, bit-exact
    assert 40_000 < len(q) < 130_000
    reg.setInputTarget(gmap)
    reg.setInputSource(q)
    reg.align()
    p = oracle.default_params("mapper")
    p.max_iterations = 50
    o = oracle.align(p, q, gmap, record_iter=-1)
    assert reg.iterations == o["iterations"]
    assert_close(reg.getFinalTransformation(), o["T"])
    idx, _ = reg.getCorrespondences()
    assert sha(idx) == sha(o["corr_idx"])
    # stand-alone exact search of the same queries against the 5 M points
    si, sd = reg.nearestKSearch1(q)
    oi, od = oracle.KdTree(gmap).nn(q)
    assert sha(si) == sha(oi) and np.array_equal(sd, od)
