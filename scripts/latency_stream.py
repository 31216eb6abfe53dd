"""Per-scan latency of the thin-cloud path (BASELINE.json configs[2]: 2-D planar 1080-point scans arriving at
40 Hz, scan-to-scan, 10 iterations): what IcpOdometer::laserCloudCallback does per message, timed on the host
around the C-ABI calls (host buffers in, result out).  Prints one JSON line; not a bench.py metric.
    python scripts/latency_stream.py [n_scans]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import registration as R, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
_, _, scans = synth.planar_stream(3, n)
reg = R.Registration(preset=R.PRESET_ODOMETER)
reg.setInputTarget(scans[0])
lat, its = [], []
for i in range(1, n):
    t0 = time.perf_counter()
    reg.setInputSource(scans[i])          # icp.setInputSource(curr_cloud_)
    reg.align()                           # icp.align + getFinalTransformation + hasConverged
    fit = reg.getFitnessScore()           # icp.getFitnessScore (icp_odometer.cpp:201)
    reg.promoteSourceToTarget()           # *prev_cloud_ = *curr_cloud_ (icp_odometer.cpp:209)
    lat.append(time.perf_counter() - t0)
    its.append(reg.iterations)
lat = np.array(lat[20:]) * 1e3
print(json.dumps({"workload": "configs[2]: 1080-pt planar scan-to-scan, 10 iterations max, stream of %d scans" % n,
                  "latency_ms_p50": float(np.percentile(lat, 50)), "latency_ms_p99": float(np.percentile(lat, 99)),
                  "latency_ms_max": float(lat.max()), "sustained_hz": float(1e3 / lat.mean()),
                  "mean_iterations": float(np.mean(its)), "required_hz": 40}))
