"""Ad-hoc GPU timing probe (not the bench): per-sweep device time of the fused ICP kernel."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icpslam_b200 import synth, registration as R

def probe(name, src, tgt, preset, reps=5, **over):
    reg = R.Registration(preset=preset, profile=1, **over)
    t0 = time.perf_counter(); reg.setInputTarget(tgt); t_tgt = time.perf_counter() - t0
    reg.setInputSource(src)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); reg.align(raise_on_fail=False); wall = time.perf_counter() - t0
        t = reg.timing()
        rec = (t.total_ms, t.nn_sweep_ms, t.nn_sweep_launches, wall * 1e3)
        best = rec if best is None or rec[0] < best[0] else best
    t0 = time.perf_counter(); fit = reg.getFitnessScore(); t_fit = time.perf_counter() - t0
    print(f"{name}: grid={reg.gridInfo()} set_target={t_tgt*1e3:.2f}ms iters={reg.iterations} "
          f"align_dev={best[0]:.3f}ms sweeps={best[1]:.3f}ms/{best[2]} -> {best[1]/max(best[2],1)*1e3:.1f}us/sweep "
          f"wall={best[3]:.3f}ms fitness={fit:.5f} ({t_fit*1e3:.2f}ms)", flush=True)

_, _, sw = synth.sweep_sequence(4, 2)
probe("C4 pair 64k/64k", sw[1], sw[0], R.PRESET_MAPPER)
for cell in (0.25, 0.35, 0.5, 0.7, 1.0):
    probe(f"C4 pair cell={cell}", sw[1], sw[0], R.PRESET_MAPPER, reps=3, grid_cell=cell)
_, _, s4 = synth.sweep_sequence(1, 2, n_beams=64, n_az=64)
probe("C1 4k/4k", s4[1], s4[0], R.PRESET_ODOMETER)
_, _, sc = synth.planar_stream(3, 2)
probe("C3 1080", sc[1], sc[0], R.PRESET_ODOMETER)
if "--map" in sys.argv:
    t0 = time.time(); m, q, Tt = synth.local_map(2); print("map gen s", time.time() - t0, m.shape)
    probe("C2 64k/500k", q, m, R.PRESET_MAPPER)
    for cell in (0.25, 0.35, 0.5):
        probe(f"C2 cell={cell}", q, m, R.PRESET_MAPPER, reps=3, grid_cell=cell)
