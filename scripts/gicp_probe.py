"""Ad-hoc GPU probe for the GICP mode: wall time per scan vs the oracle, 64k pairs and the 500k map."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import synth, registration as R
from oracle import oracle as O
O.build(); O.set_threads(O.max_threads())

def run(name, src, tgt, preset, opreset):
    reg = R.Registration(preset=preset, mode=R.MODE_GICP_BFGS)
    t0 = time.perf_counter(); reg.setInputTarget(tgt); t_t = time.perf_counter() - t0
    reg.setInputSource(src)
    t0 = time.perf_counter(); reg.align(raise_on_fail=False); t_first = time.perf_counter() - t0   # includes target covariances
    reg.setInputSource(src)
    t0 = time.perf_counter(); reg.align(raise_on_fail=False); t_next = time.perf_counter() - t0    # target covariances cached
    t0 = time.perf_counter(); o = O.align(O.default_params(opreset, O.MODE_GICP_BFGS), src, tgt); t_o = time.perf_counter() - t0
    T, To = reg.getFinalTransformation(), o["T"]
    print(f"{name}: gpu first={t_first*1e3:.1f}ms next={t_next*1e3:.1f}ms iters={reg.iterations} launches={reg.timing().kernel_launches} | "
          f"oracle {t_o*1e3:.0f}ms iters={o['iterations']} stages={ {k: round(v) for k, v in o['stages'].items()} } | "
          f"dT={np.abs(T[:3,3]-To[:3,3]).max():.2e} dR={np.abs(T[:3,:3]-To[:3,:3]).max():.2e}", flush=True)

_, _, sw = synth.sweep_sequence(4, 3)
run("GICP 64k/64k odometer", sw[1], sw[0], R.PRESET_ODOMETER, "odometer")
run("GICP 64k/64k mapper", sw[2], sw[1], R.PRESET_MAPPER, "mapper")
if "--map" in sys.argv:
    m = synth.build_local_map(2)
    qs, _ = synth.map_queries(m, 2, 0, 2)
    run("GICP 64k/500k map", qs[0], m["map"], R.PRESET_MAPPER, "mapper")
