"""GPU probe for the batched GICP mode (not the bench): scans per second of b2icp_align_batch in GICP mode on the
bench workload (32 x 64k sweeps vs the 500k map, 30 outer iterations max) and on consecutive 64k pairs."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from icpslam_b200 import registration as R, synth

map_xyzw, sweeps = bench.load_workload(0, 32)
reg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS)
t0 = time.perf_counter(); reg.setInputTarget(map_xyzw); t_tgt = time.perf_counter() - t0
out = {}
for B in (1, 8, 32):
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        rc, res = reg.alignBatch(sweeps[:B])
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    out[f"map_B{B}"] = {"scans_per_s": B / best, "ms_per_scan": 1e3 * best / B, "rc": rc,
                        "mean_outer_iterations": float(np.mean([r.iterations for r in res])),
                        "launches_total": reg.timing().kernel_launches}
_, _, sw = synth.sweep_sequence(4, 33)
reg2 = R.Registration(preset=R.PRESET_ODOMETER, mode=R.MODE_GICP_BFGS)
best = None
for rep in range(2):
    t0 = time.perf_counter()
    rc, res = reg2.alignBatch(sw[1:], [sw[0]] + [None] * 31)
    dt = time.perf_counter() - t0
    best = dt if best is None or dt < best else best
out["pairs_B32"] = {"pairs_per_s": 32 / best, "ms_per_pair": 1e3 * best / 32, "rc": rc,
                    "mean_outer_iterations": float(np.mean([r.iterations for r in res]))}
print(json.dumps(out))
