"""Timings of the steps either side of the ICP path, GPU through the C ABI with host buffers vs the
CPU oracle on the same inputs: voxel-grid filter (K8), map insertion and map nearest-neighbour gather (K9), cloud
transform.  Prints one JSON line; not a bench.py metric.      python scripts/bench_aux.py [map_points]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import registration as R, synth
from oracle import oracle as O

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rng = np.random.default_rng(7)
_, poses, sw = synth.sweep_sequence(2, 3)                      # 64k-point sweeps
raw = np.concatenate([sw[0], synth.as_xyzw(sw[0][:, :3] + rng.normal(0, 0.03, (len(sw[0]), 3)).astype(np.float32))])  # 131k "raw" points
xy = rng.uniform([-300, -200], [300, 200], (n_map, 2))
z = 2.0 * np.sin(xy[:, 0] / 30.0) * np.cos(xy[:, 1] / 40.0)
big = synth.as_xyzw(np.concatenate([xy, z[:, None]], axis=1))

def best(f, reps=5):
    f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)

reg = R.Registration(preset=R.PRESET_MAPPER)
out = {"map_points_offered": n_map, "sweep_points": len(sw[0]), "raw_points": len(raw), "host_threads": O.set_threads(O.max_threads())}
out["voxel_filter_ms"] = {"gpu": best(lambda: reg.voxelFilterCloud(raw, 0.2)), "cpu_oracle": best(lambda: O.voxel_filter(raw, 0.2), 2),
                          "points_out": len(reg.voxelFilterCloud(raw, 0.2))}
def fill():
    reg.resetMap(0.2)
    reg.addPointsToMap(big)
t0 = time.perf_counter(); fill(); t_fill = 1e3 * (time.perf_counter() - t0)
t0 = time.perf_counter(); ref = O.map_insert(None, big, 0.2); t_fill_cpu = 1e3 * (time.perf_counter() - t0)
assert reg.mapSize() == len(ref)
out["map_build_ms"] = {"gpu": best(fill, 2), "gpu_first_call": t_fill, "cpu_oracle": t_fill_cpu, "map_points": reg.mapSize()}
q = sw[1]
def ins():
    return reg.addPointsToMap(q)
t0 = time.perf_counter(); added = ins(); t_ins = 1e3 * (time.perf_counter() - t0)
t0 = time.perf_counter(); ref_added = O.map_insert(ref, q, 0.2); t_ins_cpu = 1e3 * (time.perf_counter() - t0)
assert added == len(ref_added)
out["map_insert_sweep_ms"] = {"gpu": t_ins, "cpu_oracle": t_ins_cpu, "added": added}
reg.approxNearestNeighbors(q)
out["map_nearest_sweep_ms"] = {"gpu_incl_grid_rebuild_after_insert": None, "gpu": best(lambda: reg.approxNearestNeighbors(q))}
reg.addPointsToMap(sw[2]); t0 = time.perf_counter(); reg.approxNearestNeighbors(q)
out["map_nearest_sweep_ms"]["gpu_incl_grid_rebuild_after_insert"] = 1e3 * (time.perf_counter() - t0)
T = np.eye(4); T[:3, 3] = [0.1, 0.2, 0.3]
out["transform_cloud_ms"] = {"gpu": best(lambda: reg.transformPointCloud(q, T, double=False)),
                             "cpu_oracle": best(lambda: O.transform_cloud(q, T, False), 2)}
print(json.dumps(out))
