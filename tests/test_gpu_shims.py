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


@pytest.mark.parametrize("mode", ["gicp", "p2p"])
def test_icp_odometer_shim_matches_oracle_chain(driver, oracle, tmp_path, mode):
    """IcpOdometer::laserCloudCallback over 4 sweeps: each T equals the oracle's align of (scan i, scan i-1) — GICP,
    the shims' default and what the reference instantiates (icp_odometer.cpp:188), and the point-to-point pipeline
    (B2_SHIM_MODE=p2p); the pose chain equals prev_pose o T (icp_odometer.cpp:109-113)."""
    _, _, sw = synth.sweep_sequence(7, 4, n_beams=64, n_az=128)
    env = dict(os.environ)
    env.pop("B2_SHIM_MODE", None)
    if mode == "p2p":
        env["B2_SHIM_MODE"] = "p2p"
    out = subprocess.run([driver, "odom", "0"] + _write(tmp_path, sw), check=True, capture_output=True, text=True, env=env).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert len(rows) == 4 and rows[0]["ready"] == 0
    pose = pose6dof.identity()
    p = oracle.default_params("odometer", oracle.MODE_GICP_BFGS if mode == "gicp" else oracle.MODE_P2P_SVD)
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
    env = dict(os.environ, B2_SHIM_MODE="p2p")
    out = subprocess.run([driver, "odom", "0.2"] + paths, check=True, capture_output=True, text=True, env=env).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert all(r["status"] == 0 for r in rows)
    out = subprocess.run([driver, "map", "0.2"] + paths, check=True, capture_output=True, text=True, env=env).stdout
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


def test_mapper_shim_in_reference_configuration_gicp_on_the_pcl_octree(driver, tmp_path):
    """The shims in the reference's own configuration: GICP (their default) on the PCL-compatible octree map
    (B2_SHIM_OCTREE=1).  The refined pose stays close to the odometry pose on clean data and the map grows."""
    _, _, sw = synth.sweep_sequence(8, 4, n_beams=64, n_az=128)
    paths = _write(tmp_path, sw)
    env = dict(os.environ, B2_SHIM_OCTREE="1")
    env.pop("B2_SHIM_MODE", None)
    out = subprocess.run([driver, "map", "0.2"] + paths, check=True, capture_output=True, text=True, env=env).stdout
    rows = [json.loads(l) for l in out.strip().splitlines()]
    assert rows[0]["refined"] == 0 and rows[0]["map_points"] > 1000
    assert rows[-1]["map_points"] > rows[0]["map_points"] and all(r["status"] == 0 for r in rows)
    d = np.array(rows[-1]["map_pose"][:3]) - np.array(rows[-1]["pose"][:3])
    assert np.abs(d).max() < 0.5


def test_pointcloud2_payload_to_cloud(driver, b2lib, tmp_path):
    """pcl::fromROSMsg for PointXYZ (icp_odometer.cpp:168,173) through the C ABI and through the adapter's
    fromROSMsg: a Velodyne-style payload (x, y, z, intensity f32 + ring u16 + time f32 = 22 bytes per point, fields at
    unaligned offsets) and the 16-byte {x, y, z, pad} layout; NaN returns are copied as they are."""
    rng = np.random.default_rng(0)
    n = 5000
    xyz = rng.uniform(-50, 50, (n, 3)).astype(np.float32)
    xyz[17] = np.nan
    dt = np.dtype({"names": ["i", "x", "ring", "y", "z", "t"], "formats": ["<f4", "<f4", "<u2", "<f4", "<f4", "<f4"],
                   "offsets": [0, 4, 8, 10, 14, 18], "itemsize": 22})
    msg = np.zeros(n, dt)
    msg["x"], msg["y"], msg["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    reg = b2lib.Registration()
    out = reg.fromROSMsg(msg.tobytes(), n, 1, 22, 22 * n, offsets=(4, 10, 14))
    assert np.array_equal(out[:, :3], xyz, equal_nan=True) and (out[:, 3] == 1).all()
    be = msg.astype(dt.newbyteorder(">"))                         # a big-endian sender
    assert np.array_equal(reg.fromROSMsg(be.tobytes(), n, 1, 22, 22 * n, offsets=(4, 10, 14), is_bigendian=True)[:, :3], xyz,
                          equal_nan=True)
    two_rows = reg.fromROSMsg(msg.tobytes() + b"\0" * 8, n // 2, 2, 22, 22 * (n // 2) + 4, offsets=(4, 10, 14))   # padded rows
    assert two_rows.shape == (n // 2 * 2, 4)
    with pytest.raises(b2lib.B2icpError):
        reg.fromROSMsg(msg.tobytes(), n, 1, 22, 22 * n, offsets=(4, 10, 20))      # field past the point
    # the C++ adapter (b2icp_ros_adapter.hpp): strided payload through the device, 16-byte payload through memcpy
    path = str(tmp_path / "payload.bin")
    msg[:64].tofile(path)
    j = json.loads(subprocess.run([driver, "pc2", "64", "22", "4", "10", "14", path], check=True, capture_output=True, text=True).stdout)
    got = np.array(j["xyzw"], np.float32).reshape(64, 4)
    assert j["status"] == 0 and np.array_equal(got[:, :3], xyz[:64], equal_nan=True)
    packed = np.concatenate([xyz[:64], np.zeros((64, 1), np.float32)], axis=1)
    packed.tofile(path)
    j = json.loads(subprocess.run([driver, "pc2", "64", "16", "0", "4", "8", path], check=True, capture_output=True, text=True).stdout)
    got = np.array(j["xyzw"], np.float32).reshape(64, 4)
    assert np.array_equal(got[:, :3], xyz[:64], equal_nan=True) and (got[:, 3] == 1).all()
