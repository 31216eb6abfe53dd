"""Tuning probe (not the bench): per-iteration device time of the fused sweep on the bench workload
(32 x 64k sweeps vs the 500k map, one synchronous batch) under different tuning knobs (environment variables
read at b2icp_create)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from icpslam_b200 import registration as R

configs = [dict(a.split("=") for a in c.split(",") if a) for c in sys.argv[1:]] or [{}]
map_xyzw, sweeps = bench.load_workload(0, 32)
for cfg in configs:
    for k in ("B2ICP_NO_STAGE", "B2ICP_QPT", "B2ICP_QPT_SCHED", "B2ICP_MARGIN", "B2ICP_NO_GRAPH"):
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ["B2ICP_" + k] = v
    reg = R.Registration(preset=R.PRESET_MAPPER, profile=1)
    reg.setInputTarget(map_xyzw)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        rc, res = reg.alignBatch(sweeps)
        wall = time.perf_counter() - t0
        tm = reg.timing()
        rec = (tm.total_ms, tm.nn_sweep_ms, tm.nn_sweep_launches, tm.nn_searches, wall * 1e3)
        best = rec if best is None or rec[0] < best[0] else best
    its = [r.iterations for r in res]
    print(json.dumps({"cfg": cfg, "total_ms": best[0], "sweep_ms": best[1], "launches": best[2], "searches": best[3],
                      "wall_ms": best[4], "mean_iters": float(np.mean(its))}), flush=True)
    del reg
