"""Tuning probe (not the bench): the configs[3] leg of bench.py (64 consecutive 64k pairs with getFitnessScore, host
sweeps in, results out) once warmed up — wall time per call, and with NCU=1 nothing else so that an ncu launch list of
this process shows the leg's kernels only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from icpslam_b200 import registration as R, synth

P = bench.PAIRS_PER_RANK
world_model = synth.make_world(1000 * 4)
poses = synth.trajectory(1000 * 4 + 999, P + 1)
sw = [synth.hdl64_sweep(world_model, poses[i], np.random.default_rng(1000 * 4 + i)) for i in range(P + 1)]
pinned = []
for x in sw:
    p = R.pinned_empty(x.shape)
    p[:] = x
    pinned.append(p)
srcs, tgts = pinned[1:], [pinned[0]] + [None] * (P - 1)
knobs = [dict(a.split("=") for a in c.split(",") if a) for c in sys.argv[1:]] or [{}]
for cfg in knobs:
    for k in ("B2ICP_FITNESS_RINGS",):
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ["B2ICP_" + k] = v
    reg = R.Registration(preset=R.PRESET_MAPPER)
    reps = 1 if os.environ.get("NCU") else 4
    for rep in range(reps):
        t0 = time.perf_counter()
        rc, res = reg.alignBatch(srcs, tgts, with_fitness=True)
        dt = time.perf_counter() - t0
        print(json.dumps({"cfg": cfg, "rep": rep, "rc": rc, "ms": 1e3 * dt, "pairs_per_s": P / dt,
                          "mean_iterations": float(np.mean([r.iterations for r in res])),
                          "fitness_sum": float(np.sum([r.fitness for r in res])),
                          "launches": reg.timing().kernel_launches}), flush=True)
    del reg
