"""GPU tests of the C++ host side (IcpOdometer / OctreeMapper shims over the C ABI) and of the
sharded replay helper, against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from icpslam_b200 import pose6dof, replay, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(b2lib, tmp_path_factory):
    from icpslam_b200 import build as B
    exe = str(tmp_path_factory.mktemp("shim") / "shim_driver")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
                    "-L", B.LIB_DIR, "-lb2icp", f"-Wl,-rpath,{B.LIB_DIR}"], check=True)
    return exe


def _write(tmp_path, clouds):
    paths = []
    for i, c in enumerate(clouds):
        p = str(tmp_path / f"scan{i}.bin")
        np.ascontiguousarray(c, np.float32).tofile(p)
        paths.append(p)
    return paths


def test_icp_odometer_shim_matches_oracle_chain(driver, oracle, tmp_path):
    """IcpOdometer::laserCloudCallback over 4 sweeps: each T equals the oracle's P2P align of
    (scan i, scan i-1); the pose chain equals prev_pose o T (icp_odometer.cpp:109-113)."""
    _, _, sw = synth.sweep_sequence(7, 4, n_beams=64, n_az=128)
    out = subprocess.run([driver, "odom", "0"] + _write(tmp_path, sw), check=True, capture_output=True, text=True).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert len(rows) == 4 and rows[0]["ready"] == 0
    pose = pose6dof.identity()
    p = oracle.default_params("odometer")
    for i in range(1, 4):
        o = oracle.align(p, sw[i], sw[i - 1])
        T = np.array(rows[i]["T"]).reshape(4, 4)
        assert rows[i]["status"] == 0 and rows[i]["iterations"] == o["iterations"]
        assert np.abs(T[:3, 3] - o["T"][:3, 3]).max() <= 1e-4 and np.abs(T[:3, :3] - o["T"][:3, :3]).max() <= 1e-4
        ofit = oracle.fitness(sw[i], sw[i - 1], T.astype(np.float32))
        assert abs(rows[i]["fitness"] - ofit) <= 1e-9 * max(1, ofit)
        if rows[i]["converged"] and rows[i]["fitness"] < 20:
            pose = pose6dof.compose(pose, pose6dof.from_matrix(T))
        assert np.abs(np.array(rows[i]["pose"]) - pose).max() < 1e-12
    assert rows[3]["ready"] == 1


def test_voxel_filter_and_mapper_shim_run(driver, tmp_path):
    """voxelFilterCloud (leaf 0.2) in front of the odometer, and refineTransformAndGrowMap growing a
    one-point-per-voxel map: the refined pose must stay close to the odometry pose on clean data."""
    _, _, sw = synth.sweep_sequence(8, 4, n_beams=64, n_az=256)
    paths = _write(tmp_path, sw)
    out = subprocess.run([driver, "odom", "0.2"] + paths, check=True, capture_output=True, text=True).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert all(r["status"] == 0 for r in rows)
    out = subprocess.run([driver, "map", "0.2"] + paths, check=True, capture_output=True, text=True).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert rows[0]["refined"] == 0 and rows[0]["map_points"] > 1000       # first scan seeds the map
    assert all(r["refined"] == 1 for r in rows[1:])
    assert rows[-1]["map_points"] > rows[0]["map_points"]
    d = np.array(rows[-1]["map_pose"][:3]) - np.array(rows[-1]["pose"][:3])
    assert np.linalg.norm(d) < 0.3


def test_replay_pairs_single_rank(b2lib, oracle):
    _, _, sw = synth.sweep_sequence(9, 5, n_beams=32, n_az=128)
    reg = b2lib.Registration()
    records, poses = replay.replay_pairs(sw, reg)
    assert records.shape == (4, replay.RECORD) and poses.shape == (5, 7)
    p = oracle.default_params("odometer")
    pose = pose6dof.identity()
    for i in range(4):
        o = oracle.align(p, sw[i + 1], sw[i])
        assert np.abs(records[i, :16].reshape(4, 4) - o["T"]).max() <= 1e-4
        pose = pose6dof.compose(pose, pose6dof.from_matrix(records[i, :16]))
    assert np.abs(poses[-1] - pose).max() < 1e-12
