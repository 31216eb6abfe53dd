"""CPU tests of the host-side logic above the C ABI: Pose6DOF mirror, scan sharding, the per-batch
gather (world_size 2 over gloo) and the shims' compile check."""
import os
import subprocess
import sys

import numpy as np
import pytest

from icpslam_b200 import pose6dof, replay, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pose6dof_mirror_equals_oracle(oracle):
    rng = np.random.default_rng(5)
    for _ in range(20):
        Ta, Tb = synth.random_rigid(rng, 5, 3), synth.random_rigid(rng, 5, 3)
        a, b = pose6dof.from_matrix(Ta), pose6dof.from_matrix(Tb)
        assert np.abs(a - oracle.pose_from_matrix(Ta)).max() < 1e-14
        assert np.abs(pose6dof.compose(a, b) - oracle.pose_compose(a, b)).max() < 1e-14
        assert np.abs(pose6dof.inverse(a) - oracle.pose_inverse(a)).max() < 1e-14
        assert np.abs(pose6dof.to_matrix(pose6dof.compose(a, b)) - Ta @ Tb).max() < 1e-12


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 511, 512):
        for world in (1, 2, 3, 4, 8):
            blocks = [replay.shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_compose_odometry_skips_rejected_scans():
    """A rejected scan keeps the previous pose AND the previous target (reference icp_odometer.cpp:201-209)."""
    T = np.eye(4)
    T[0, 3] = 1.0
    recs = np.zeros((3, replay.RECORD))
    recs[:, :16] = T.reshape(-1)
    recs[:, 16] = [1, 0, 1]                      # scan 2 did not converge: dropped, pose kept
    recs[:, 20] = np.nan                         # no fitness asked for
    recs[:, 21] = [0, 1, 1]                      # pair 2 was (re-)registered against sweep 1, the last accepted one
    poses = replay.compose_odometry(recs)
    assert poses[:, 0].tolist() == [0, 1, 1, 2]
    recs[:, 16] = 1
    recs[:, 20] = [0.1, 0.1, 25.0]               # fitness >= 20 rejected
    recs[:, 21] = [0, 1, 2]
    poses = replay.compose_odometry(recs)
    assert poses[:, 0].tolist() == [0, 1, 2, 2]
    # a table that still registers pair 2 against the REJECTED sweep 2 must not be composed silently
    recs[:, 20] = [0.1, 25.0, 0.1]
    with pytest.raises(ValueError):
        replay.compose_odometry(recs)


class _FakeResult:
    def __init__(self, dx, converged=1, fitness=0.1):
        T = np.eye(4)
        T[0, 3] = dx
        self.T, self.converged, self.iterations, self.n_corr_last, self.mse_last, self.fitness = T.reshape(-1), converged, 3, 100, 0.01, fitness


class _FakeRegistration:
    """Sweeps are 1-point 'clouds' holding their x position: the 'registration' returns the x offset between the
    two, and rejects (fitness 99) any pair whose source is marked bad."""
    def __init__(self):
        self.calls = []

    def alignBatch(self, sources, targets, with_fitness=False):
        assert with_fitness
        out, prev = [], None
        for s, t in zip(sources, targets):
            t = prev if t is None else t
            self.calls.append((float(t[0, 0]), float(s[0, 0])))
            out.append(_FakeResult(s[0, 0] - t[0, 0], fitness=99.0 if s[0, 3] < 0 else 0.1))
            prev = s
        return 0, out


def test_replay_reregisters_against_the_last_accepted_sweep():
    """after a rejected scan the reference keeps prev_cloud_, so the NEXT scan is registered against
    the older cloud.  The batch registers every sweep against its predecessor; fixup_rejected repairs the pairs
    behind a rejection so that no motion is lost."""
    xs = [0.0, 1.0, 2.5, 3.0, 4.0]
    sweeps = [np.array([[x, 0, 0, 1.0]], np.float32) for x in xs]
    sweeps[2][0, 3] = -1.0                       # sweep 2 is garbage: every registration OF it is rejected
    reg = _FakeRegistration()
    records, poses = replay.replay_pairs(sweeps, reg)
    assert [int(t) for t in records[:, 21]] == [0, 1, 1, 3]       # pair 2 (sweep 3) went back to sweep 1
    assert reg.calls[-1] == (1.0, 3.0)
    assert poses[:, 0].tolist() == [0.0, 1.0, 1.0, 3.0, 4.0]      # the motion across the rejected sweep is kept


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 7
    lo, hi = replay.shard_range(n_total, world, rank)
    local = np.zeros((hi - lo, replay.RECORD))
    for i in range(lo, hi):
        T = np.eye(4)
        T[0, 3] = i + 1
        local[i - lo, :16] = T.reshape(-1)
        local[i - lo, 16:] = (1, 5 + i, 100 + i, 0.5, 0.25, i)
    table = replay.gather_records(local, n_total)
    poses = replay.compose_odometry(table)
    q.put((rank, table[:, 3].tolist(), table[:, 17].tolist(), poses[-1, 0]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_of_transforms_world2_gloo():
    """The only exchange step of the path: per-scan records to every rank, in scan order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tx, iters, last_x in got:
        assert tx == [1, 2, 3, 4, 5, 6, 7]
        assert iters == [5, 6, 7, 8, 9, 10, 11]
        assert last_x == sum(range(1, 8))


def test_shims_compile_and_pose_selfcheck(b2lib, tmp_path):
    """The C++ shims (IcpOdometer / OctreeMapper / Pose6DOF) build against include/b2icp.h and link to
    libb2icp.so; their Pose6DOF algebra needs no GPU and is checked here."""
    from icpslam_b200 import build as B
    exe = str(tmp_path / "shim_driver")
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"),
           "-L", B.LIB_DIR, "-lb2icp", f"-Wl,-rpath,{B.LIB_DIR}"]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe, "pose"], check=True, capture_output=True, text=True).stdout
    import json
    j = json.loads(out)
    T = np.array([[0, -1, 0, 1], [1, 0, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1.0]])
    assert np.abs(np.array(j["a"]) - pose6dof.from_matrix(T)).max() < 1e-15
    assert np.abs(np.array(j["ab"]) - pose6dof.compose(pose6dof.from_matrix(T), pose6dof.from_matrix(T))).max() < 1e-15
    assert np.abs(np.array(j["ident"]) - pose6dof.identity()).max() < 1e-15
    import torch
    if not torch.cuda.is_available():             # no CPU fallback behind the shims either
        r = subprocess.run([exe, "odom", "0.2", "/dev/null", "/dev/null"], capture_output=True, text=True)
        assert r.returncode == 3 and "CUDA" in r.stderr


@pytest.mark.parametrize("flags", [[], ["-DB2_FIBER_UCONTEXT"]], ids=["register-switch", "ucontext"])
def test_gicp_fibers_interleave_without_disturbing_each_other(tmp_path, flags):
    """The batched GICP host loop runs every scan's blocking recursion on its own fiber (icpslam_b200/csrc/fiber.h): 32
    fibers doing floating-point and libc work between different numbers of yields, interleaved by a coordinator, end
    with bitwise the results of the same work done straight — for the six-register switch and for its fallback."""
    exe = str(tmp_path / "fiber_test")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", *flags, "-o", exe, os.path.join(ROOT, "tests", "cpp", "fiber_test.cpp")],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
    assert ("ucontext" in r.stdout) == bool(flags) or os.uname().machine != "x86_64"


def test_ros_adapter_compiles_against_stub_headers(b2lib, tmp_path):
    """The ROS-facing half of shims/b2icp_ros_adapter.hpp (::IcpOdometer / ::OctreeMapper with the reference's
    signatures, include/icpslam/icp_odometer.h:30-58, octree_mapper.h:23-49) only exists behind __has_include(<ros/ros.h>)
    and this image has no ROS: it is compiled here against minimal stand-in headers (tests/cpp/ros_stubs) with every
    adapter method referenced through the reference's argument types, and linked against libb2icp.so.  Without a GPU
    the constructed odometer must fail loudly (no CPU path behind the adapter either)."""
    from icpslam_b200 import build as B
    exe = str(tmp_path / "ros_adapter")
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "tests", "cpp", "ros_stubs"), "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "ros_adapter_compile.cpp"), "-L", B.LIB_DIR, "-lb2icp", f"-Wl,-rpath,{B.LIB_DIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([exe], capture_output=True, text=True).stdout.strip() == "compiled"
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "run"], capture_output=True, text=True)
        assert r.returncode == 3 and "CUDA" in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle timed on the host cores, no GPU, nothing read from /root/reference)
    prints one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "icp_scans_per_sec_64k_sweeps_30_iters"
    assert line["unit"] == "scans/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
