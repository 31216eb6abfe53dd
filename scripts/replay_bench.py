"""BASELINE.json configs[3] on one GPU: consecutive 64k-point sweeps, pair i registers against sweep i-1 (30 iterations
max), through icpslam_b200/replay.py (b2icp_align_batch in consecutive mode: every source is uploaded once and indexed
in place as the next pair's target).  Prints one JSON line; not a bench.py metric.   python scripts/replay_bench.py [n]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import registration as R, replay, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
_, poses, sw = synth.sweep_sequence(4, n)
pinned = []
for s in sw:
    p = R.pinned_empty(s.shape); p[:] = s; pinned.append(p)
reg = R.Registration(preset=R.PRESET_MAPPER)
replay.replay_pairs(pinned, reg)                      # warm-up: buffers, grids
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    records, traj = replay.replay_pairs(pinned, reg)
    ts.append(time.perf_counter() - t0)
dt = min(ts)
true_rel = [np.linalg.inv(poses[i]) @ poses[i + 1] for i in range(n - 1)]
err = [np.abs(records[i, :16].reshape(4, 4)[:3, 3] - true_rel[i][:3, 3]).max() for i in range(n - 1)]
print(json.dumps({"workload": "configs[3] on 1 GPU: %d consecutive 64k-pt sweeps, %d pairs, 30 iterations max" % (n, n - 1),
                  "pairs_per_s": (n - 1) / dt, "ms_per_pair": 1e3 * dt / (n - 1), "mean_iterations": float(records[:, 17].mean()),
                  "converged": int(records[:, 16].sum()), "max_translation_error_vs_ground_truth_m": float(max(err))}))
