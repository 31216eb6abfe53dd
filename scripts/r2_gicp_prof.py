"""Tuning probe: one GICP batch of 8 sweeps vs the 500k map, with host timestamps of the phases."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from icpslam_b200 import registration as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
map_xyzw, sweeps = bench.load_workload(0, 32)
reg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS)
reg.setInputTarget(map_xyzw)
for rep in range(3):
    t0 = time.perf_counter()
    rc, res = reg.alignBatch(sweeps[:B])
    print("rep", rep, "ms", 1e3 * (time.perf_counter() - t0), "iters", [r.iterations for r in res], flush=True)
