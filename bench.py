#!/usr/bin/env python
"""bench.py — ICP scans/sec on BASELINE.json configs[1]:
64k-pt synthetic Velodyne HDL-64 sweeps registered against a 500k-pt accumulated local map,
30 ICP iterations max (point-to-point pipeline of the north_star), one B200 per rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

A "step" is one pass of the hot path over one batch of B sweeps (b2icp_align_batch_submit / _wait against the
resident map; one fused sweep launch per ICP iteration for the whole batch).  Steps rotate through --sets (4)
DISTINCT sets of B sweeps.  Weak scaling: every rank owns its own sets; the map is replicated.  The only exchange
of the path is the gather of the per-scan rigid transforms: the library appends every
scan's 96-byte record to a device buffer (b2icp_set_record_sink) and ONE ncclAllGather on a side stream, after the
last step and inside the timed region, hands all K*B records of every rank to every rank — no per-step barrier.

  value     scans/s with the sweeps already resident in HBM (device pointers through the C ABI)
  e2e       scans/s through the same C ABI with pinned HOST buffers: H2D of every sweep and D2H of the
            results inside the timed region
  roofline  fused sweep kernel: algorithmic bytes / CUDA-event launch time, vs the measured HBM peak; nested in it:
            nn_search (the stand-alone search kernel), pairs (BASELINE configs[3]: consecutive 64k-pt sweeps, each
            registered against its predecessor, 64 pairs per rank + the gather of the records), gicp (the same sweeps
            through the reference's own estimator), details
  parity    the GPU transforms / iteration counts of the --cpu-sample sweeps against the CPU oracle's (same run)
  cpu_baseline  the CPU oracle (kind "port": the reference's PCL path cannot be built here) timed on
            this box's host cores on a bounded sample of the same workload

`--impl reference` times the oracle alone (all host threads) on the same workload and prints the
same line with "impl": "reference".  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from icpslam_b200 import synth  # noqa: E402

CONFIG_ID = 2          # seeds = 1000 * config + index
N_MAP = 500_000
N_SWEEP = 65_536
MAX_ITERS = 30
METRIC = "icp_scans_per_sec_64k_sweeps_30_iters"
UNIT = "scans/s"
STATE_BYTES = 192      # sizeof(IcpState): what comes back per scan
N_SETS = 4             # distinct sets of sweeps the steps rotate through
PAIRS_PER_RANK = 64    # configs[3] leg: consecutive pairs per rank (8 ranks = the 512 sweeps of BASELINE configs[3])
GICP_CALLERS = 4       # host threads (one handle each) of the concurrent GICP leg


_T0 = time.time()


def log(*a):
    print(f"[{time.time() - _T0:6.1f}s]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def load_workload(first: int, count: int, cache_dir: str = "/tmp/b2icp_bench_cache"):
    """(map[500000,4], [count sweeps already in the map frame up to a small odometry error])."""
    return prepare_workloads([(first, 0)], count, cache_dir)[(first, 0)]


def prepare_workloads(keys, count: int, cache_dir: str = "/tmp/b2icp_bench_cache") -> dict:
    """{(first, variant): (map, sweeps)}: `count` sweeps from pose index `first` on; variant v > 0 re-draws the range
    noise and the odometry error of the same poses.  The map is generated once and all of it is cached under /tmp
    as .npy so that back-to-back runs on one box skip the ray casting."""
    os.makedirs(cache_dir, exist_ok=True)
    mpath = os.path.join(cache_dir, f"map_c{CONFIG_ID}_{N_MAP}.npy")
    qpaths = {k: os.path.join(cache_dir, f"q_c{CONFIG_ID}_{k[0]}_{count}" + (f"_v{k[1]}" if k[1] else "") + ".npy") for k in keys}
    out, m = {}, None
    for k in keys:
        if os.path.exists(mpath) and os.path.exists(qpaths[k]):
            out[k] = (np.load(mpath), list(np.load(qpaths[k])))
            continue
        t0 = time.time()
        if m is None:
            m = synth.build_local_map(CONFIG_ID, n_points=N_MAP)
        qs, _ = synth.map_queries(m, CONFIG_ID, k[0], count, variant=k[1])
        log(f"[bench] generated workload (first sweep {k[0]}, variant {k[1]}, {count} sweeps) in {time.time() - t0:.1f}s")
        try:
            tmp = os.path.join(cache_dir, f"tmp_{os.getpid()}.npy")
            if not os.path.exists(mpath):
                np.save(tmp, m["map"])
                os.replace(tmp, mpath)
            np.save(tmp, np.stack(qs))
            os.replace(tmp, qpaths[k])
        except OSError:
            pass
        out[k] = (m["map"], qs)
    return out


def touched_target_points(map_xyzw, queries, cell, origin, dims, radius):
    """N_t' of SURVEY.md §8d: map points lying in grid cells within `radius` of any query's cell."""
    dims = np.asarray(dims, dtype=np.int64)
    R = int(np.ceil(radius / cell))
    q = np.concatenate([x[:, :3] for x in queries]).astype(np.float64)
    qc = np.clip(np.floor((q - origin) / cell).astype(np.int64), 0, dims - 1)
    lin = np.unique((qc[:, 2] * dims[1] + qc[:, 1]) * dims[0] + qc[:, 0])
    cz, rem = np.divmod(lin, dims[0] * dims[1])
    cy, cx = np.divmod(rem, dims[0])
    touched = np.zeros(int(dims.prod()), dtype=bool)
    for dz in range(-R, R + 1):
        z = cz + dz
        okz = (z >= 0) & (z < dims[2])
        for dy in range(-R, R + 1):
            y = cy + dy
            ok = okz & (y >= 0) & (y < dims[1])
            base = (z[ok] * dims[1] + y[ok]) * dims[0]
            x0 = np.clip(cx[ok] - R, 0, dims[0] - 1)
            x1 = np.clip(cx[ok] + R, 0, dims[0] - 1)
            for dx in range(0, 2 * R + 1):
                x = np.minimum(x0 + dx, x1)
                touched[base + x] = True
    m = map_xyzw[:, :3].astype(np.float64)
    mc = np.clip(np.floor((m - origin) / cell).astype(np.int64), 0, dims - 1)
    ml = (mc[:, 2] * dims[1] + mc[:, 1]) * dims[0] + mc[:, 0]
    return int(touched[ml].sum())


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("B2_BENCH_CLOCK_MS", "200"), "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle legs (the only places bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_scans_per_sec(map_xyzw, sweeps, threads: int):
    """P2P align of each sweep vs the map with the CPU oracle; the k-d tree build is excluded from the
    time like the resident grid is excluded on the GPU side.  Returns (scans/s, threads, iterations)."""
    from oracle import oracle as O
    O.build()
    threads = O.set_threads(threads)
    p = O.default_params("mapper")
    busy_ms, iters, results = 0.0, [], []
    for s in sweeps:
        r = O.align(p, s, map_xyzw)
        busy_ms += r["stages"]["total"] - r["stages"]["build"]
        iters.append(r["iterations"])
        results.append(r)
    oracle_scans_per_sec.results = results
    return len(sweeps) / (busy_ms * 1e-3), threads, iters


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = O.max_threads()
    per_step = 2
    map_xyzw, sweeps = load_workload(0, max(args.batch, per_step))
    t_steps = []
    for k in range(args.warmup + args.steps):
        sample = [sweeps[(k * per_step + j) % len(sweeps)] for j in range(per_step)]
        t0 = time.perf_counter()
        sps, used, iters = oracle_scans_per_sec(map_xyzw, sample, threads)
        if k >= args.warmup:
            t_steps.append(per_step / sps)
    total = float(np.sum(t_steps))
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"each step is a bounded sample of the workload: {per_step} of the batch's sweeps per step "
                                   f"x {args.steps} steps vs the 500k map, k-d tree build excluded (map resident), OpenMP "
                                   f"over queries; scans/s of the serial oracle does not depend on the batch size"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(batch: int, world: int) -> dict:
    """The SAME dict in both arms (the driver compares them)."""
    c = {
        "workload": "BASELINE configs[1]: 64k-pt synthetic HDL-64 sweep vs 500k-pt accumulated local map, "
                    "30 ICP iterations max, point-to-point (north_star pipeline)",
        "sweep_points": N_SWEEP, "map_points": N_MAP, "max_iterations": MAX_ITERS, "transformation_epsilon": 1e-6,
        "max_correspondence_distance": 1.0, "batch_per_gpu": batch, "global_batch": batch * world,
        "sets": N_SETS,
        "sharding": "scans sharded across ranks, map replicated; one NCCL all_gather of the per-scan records per run "
                    "(device record sink, side stream), inside the timed region",
        "l2": "a 256 MiB write is enqueued between step submissions (in the streamed legs it runs next to the batches in "
              "flight); independently of it the per-step working set (150 MB of per-query state per 32-sweep step, up to "
              "8 steps in flight, steps rotating through 4 distinct sets of sweeps) is larger than the 126 MB L2",
    }
    return c


# ------------------------------------------------------------------------------------------------
def rot_angle(Ra, Rb):
    R = Ra.T @ Rb
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arcsin(min(1.0, float(np.linalg.norm(v)))))


def pairs_leg(R, replay, dev, rank, world, local_rank, reps=2):
    """BASELINE configs[3]: consecutive 64k-pt sweeps, sweep i registered against sweep i-1 (30 iterations max,
    getFitnessScore for the reference's accept test), PAIRS_PER_RANK pairs per rank, then the gather of the
    records (one collective) and the serial pose composition.  Host sweeps in, records out: end to end."""
    import torch
    import torch.distributed as dist
    # The 120 m scene holds ~250 consecutive sweeps of the drive: rank r replays the block of 65 sweeps that starts at
    # sweep (r mod 3) * 64, with its own range noise (seed offset r), so that 8 ranks hold 8 different recordings.
    first = (rank % 3) * PAIRS_PER_RANK
    cache = f"/tmp/b2icp_bench_cache/pairs_c4_{first}_{PAIRS_PER_RANK + 1}_r{rank}.npy"
    if os.path.exists(cache):
        sw = list(np.load(cache))
    else:
        world_model = synth.make_world(1000 * 4)
        poses = synth.trajectory(1000 * 4 + 999, first + PAIRS_PER_RANK + 1)
        sw = [synth.hdl64_sweep(world_model, poses[i], np.random.default_rng(1000 * 4 + i + 1_000_000 * rank))
              for i in range(first, first + PAIRS_PER_RANK + 1)]
        try:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.save(cache, np.stack(sw))
        except OSError:
            pass
    pinned = []
    for x in sw:
        p = R.pinned_empty(x.shape)
        p[:] = x
        pinned.append(p)
    reg = R.Registration(preset=R.PRESET_MAPPER, device=local_rank)
    n_total = PAIRS_PER_RANK * world

    def once():
        srcs, tgts = pinned[1:], [pinned[0]] + [None] * (PAIRS_PER_RANK - 1)
        rc, res = reg.alignBatch(srcs, tgts, with_fitness=True)
        if rc in replay.HARD_ERRORS:
            raise RuntimeError(f"pairs leg: b2icp_align_batch rc={rc}")
        local = replay.pack_results(res, rank * PAIRS_PER_RANK)
        return replay.gather_records(local, n_total, dev), res

    once()  # buffers, grids
    best = None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        records, res = once()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt[0]) if best is None else min(best, float(dt[0]))
    acc = sum(1 for r in records if replay.accepted(r))
    return {"workload": "BASELINE configs[3]: consecutive 64k-pt sweeps, pair i = sweep i vs sweep i-1, 30 iterations max, "
                        "getFitnessScore per pair; host sweeps in, gathered records out",
            "pairs_per_rank": PAIRS_PER_RANK, "pairs": n_total, "pairs_per_s": n_total / best, "ms_total": 1e3 * best,
            "mean_iterations": float(np.mean(records[:, 17])), "accepted": int(acc),
            "collective": "one all_gather of 22-double records per replay" if world > 1 else "none (1 rank)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="sweeps per GPU per step")
    ap.add_argument("--impl", default="b2icp", choices=["b2icp", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=4, help="sweeps timed on the CPU oracle and parity-checked (rank 0, N=1)")
    ap.add_argument("--in-flight", type=int, default=8, help="streamed batches in flight (1..8)")
    ap.add_argument("--grid-cell", type=float, default=0.0, help="neighbour-grid cell edge in metres (0 = auto); tuning only")
    ap.add_argument("--no-pairs", action="store_true", help="skip the configs[3] leg")
    ap.add_argument("--no-gicp", action="store_true", help="skip the GICP-mode leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2icp" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from icpslam_b200 import build as B
    from icpslam_b200 import registration as R
    from icpslam_b200 import replay

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libb2icp.so has no CPU fallback")
    B.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    Bn = args.batch
    # Every rank casts its own sweeps (in parallel).  All ranks see the SAME B poses (the first B after the map) and
    # differ in the range noise and odometry error drawn for them: rank r owns noise variants r*N_SETS .. r*N_SETS+3.
    # Weak scaling needs equal work per rank: in round 1 rank r took poses r*B .. r*B+B-1, which lie farther and
    # farther from the mapped area and need more iterations (profiles/r02_scaling.md) — the slowest rank then set
    # the step time and the curve read as a scaling loss.
    keys = [(0, rank * N_SETS + v) for v in range(N_SETS)]
    loaded = prepare_workloads(keys, Bn)
    map_xyzw = loaded[keys[0]][0]
    sets = [loaded[k][1] for k in keys]
    sweeps = sets[0]

    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev)  # the gather's stream: nothing else of the step waits on it
    reg = R.Registration(preset=R.PRESET_MAPPER, device=local_rank, profile=1, grid_cell=args.grid_cell)
    reg.setStream(stream.cuda_stream)
    reg.setInputTarget(map_xyzw)
    grid = reg.gridInfo()

    d_sets = [[torch.from_numpy(x).to(dev) for x in st] for st in sets]
    d_ptrs = [[t.data_ptr() for t in st] for st in d_sets]
    n_src = [N_SWEEP] * Bn
    h_sets = []
    for st in sets:
        hs = []
        for x in st:
            p = R.pinned_empty((N_SWEEP, 4))
            p[:] = x
            hs.append(p)
        h_sets.append(hs)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def timed_sync(steps, warmup, collect):
        """One synchronous b2icp_align_batch_device per step with CUDA events around every sweep launch inside the
        library (params.profile = 1): the per-launch durations the roofline needs.  Not the headline numbers."""
        for k in range(warmup):
            reg.alignBatchDevice(d_ptrs[k % N_SETS], n_src)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        first_results = None
        for k in range(steps):
            flush.zero_()  # L2 flush between timed steps, outside the per-step events
            ev[k][0].record(stream)
            rc, res = reg.alignBatchDevice(d_ptrs[k % N_SETS], n_src)
            if rc:
                raise RuntimeError(f"b2icp_align_batch_device rc={rc}")
            ev[k][1].record(stream)
            collect(res)
            if k == 0:
                first_results = res
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) * 1e-3, first_results

    sweep_ms, sweep_launches, iters_hist, searches = [], [], [], []

    def collect(res):
        tm = reg.timing()
        sweep_ms.append(tm.nn_sweep_ms)
        sweep_launches.append(tm.nn_sweep_launches)
        searches.append(tm.nn_searches)
        iters_hist.append([r.iterations for r in res])

    log("[bench] workload resident; profiled synchronous pass")
    prof_dev_s, gpu_first = timed_sync(args.steps, args.warmup, collect)
    log("[bench] streamed legs")

    # The two timed legs use the streaming form of the batch call (b2icp_align_batch_submit[_device] / _wait): up to
    # --in-flight steps are submitted before the oldest is waited for, each on its own stream.  Every step's copies
    # (H2D of its sweeps, D2H of its results), the L2 flush between steps and — with more than one rank — the ONE
    # gather of all records at the end are inside the timed region: one event pair around all K steps.
    host = {"submit_s": 0.0, "wait_s": 0.0}
    no_flush = os.environ.get("BENCH_NO_FLUSH") is not None  # diagnostics only

    def run_streamed(steps, submit):
        submitted = done = 0
        while done < steps:
            t0 = time.perf_counter()
            while submitted < steps and submitted - done < args.in_flight:
                if submitted and not no_flush:
                    flush.zero_()
                if submit(submitted):
                    raise RuntimeError("b2icp_align_batch_submit failed")
                submitted += 1
            t1 = time.perf_counter()
            rc, res = reg.alignBatchWait()
            if rc:
                raise RuntimeError(f"b2icp_align_batch_wait rc={rc}")
            host["submit_s"] += t1 - t0
            host["wait_s"] += time.perf_counter() - t1
            done += 1

    rec_words = R.RECORD_DTYPE.itemsize // 4
    cap = max(args.steps, args.warmup, 17) * Bn
    sink = torch.zeros((cap, rec_words), dtype=torch.int32, device=dev)
    gathered = torch.empty((world * cap, rec_words), dtype=torch.int32, device=dev) if world > 1 else None
    gather_ev = torch.cuda.Event(enable_timing=False)

    def timed_streamed(submit, steps, warmup):
        reg.setRecordSink(None, 0)
        # every one of the library's 8 slot sets allocates its buffers the first time it is used and captures its
        # loop into a CUDA graph the second time: both belong to the warm-up, not to the timed steps
        run_streamed(max(warmup, 17), submit)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        reg.setRecordSink(sink.data_ptr(), cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = reg.timing().kernel_launches
        host["submit_s"] = host["wait_s"] = 0.0
        e0.record(stream)
        run_streamed(steps, submit)
        timed_streamed.host = dict(host)
        if world > 1:  # every batch has been waited for: its records are in `sink`
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(gathered, sink)
                gather_ev.record(side)
            stream.wait_event(gather_ev)
        e1.record(stream)
        timed_streamed.launches = reg.timing().kernel_launches - n0
        torch.cuda.synchronize()
        reg.setRecordSink(None, 0)
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    streamed = Bn <= 32
    if not streamed:
        raise SystemExit("bench.py: --batch must be <= 32 (one streamed batch per step)")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_s = timed_streamed(lambda k: reg.alignBatchSubmitDevice(d_ptrs[k % N_SETS], n_src), args.steps, args.warmup)
    launches = timed_streamed.launches  # the launches of the K timed steps only
    host_resident = timed_streamed.host
    log(f"[bench] rank {rank}: resident leg {1e3 * dev_s / args.steps:.3f} ms/step; host per step: submit "
        f"{1e3 * host_resident['submit_s'] / args.steps:.3f} ms, wait {1e3 * host_resident['wait_s'] / args.steps:.3f} ms")
    clocks = sampler.stop() if rank == 0 else None
    value = world * Bn * args.steps / dev_s
    # the records of the resident leg as the ranks exchanged them (rank 0 checks them against what _wait returned)
    rec_ok = None
    if world > 1 and rank == 0:
        got = gathered.cpu().numpy().view(R.RECORD_DTYPE).reshape(world, cap)
        rec_ok = bool(all(int(got[r, args.steps * Bn - 1]["iterations"]) > 0 for r in range(world)))

    e2e_dev_s = timed_streamed(lambda k: reg.alignBatchSubmit(h_sets[k % N_SETS]), args.steps, 3)
    api = f"b2icp_align_batch_submit[_device] / b2icp_align_batch_wait ({args.in_flight} batches in flight)"
    e2e_value = world * Bn * args.steps / e2e_dev_s

    # ---- the reference's own estimator on the same workload (GICP, icp_odometer.cpp:188 / octree_mapper.cpp:104): one
    # b2icp_align_batch of the B sweeps of set 0 in GICP mode (scans advance in lockstep rounds), host sweeps in
    gicp = None
    if rank == 0 and not args.no_gicp:
        greg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS, device=local_rank)
        greg.setInputTarget(map_xyzw)
        greg.alignBatch(h_sets[0])                     # target covariances (cached with the grid), buffers
        best, gres = None, None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            grc, gres = greg.alignBatch(h_sets[0])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        gicp = {"workload": "the same 32 sweeps vs the 500k map, pcl::GeneralizedIterativeClosestPoint restatement (k = 20 "
                            "covariances, Mahalanobis, BFGS), 30 outer iterations max; host sweeps in, results out",
                "scans_per_s": Bn / best, "ms_per_scan": 1e3 * best / Bn, "rc": int(grc),
                "mean_outer_iterations": float(np.mean([r.iterations for r in gres])),
                "converged": int(sum(r.converged for r in gres))}
        # the same call from GICP_CALLERS host threads at once, each with its own handle and its own set of 32 sweeps
        # (the four noise variants): a batch's rounds leave the device idle between launches and its set-up leaves the
        # host idle, so concurrent callers fill both — the counterpart of the 8 batches in flight of the main leg
        try:  # an auxiliary record: a failure here must not cost the bench line
            handles = [greg] + [R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS, device=local_rank)
                                for _ in range(GICP_CALLERS - 1)]
            for g in handles[1:]:
                g.setInputTarget(map_xyzw)
            outs = [None] * GICP_CALLERS

            def caller(k):
                torch.cuda.set_device(local_rank)
                outs[k] = handles[k].alignBatch(h_sets[k % N_SETS])

            def all_callers():
                ths = [threading.Thread(target=caller, args=(k,)) for k in range(GICP_CALLERS)]
                t0 = time.perf_counter()
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
                if any(o is None for o in outs):
                    raise RuntimeError("a concurrent GICP caller did not return")
                return time.perf_counter() - t0

            all_callers()                              # buffers and target covariances of the extra handles
            bestc = min(all_callers() for _ in range(2))
            gicp["concurrent"] = {"callers": GICP_CALLERS, "scans": GICP_CALLERS * Bn,
                                  "scans_per_s": GICP_CALLERS * Bn / bestc, "ms_total": 1e3 * bestc,
                                  "rc": [int(o[0]) for o in outs],
                                  "converged": int(sum(r.converged for o in outs for r in o[1]))}
            del handles
        except Exception as e:  # noqa: BLE001
            gicp["concurrent"] = {"error": f"{type(e).__name__}: {e}"}
        del greg
    pairs = None
    if not args.no_pairs:
        log("[bench] configs[3] leg")
        pairs = pairs_leg(R, replay, dev, rank, world, local_rank)
    log("[bench] roofline bookkeeping, stand-alone search, CPU sample")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused sweep) -------------------------------------------
    origin = map_xyzw[:, :3].min(axis=0).astype(np.float64)
    nt_touched = touched_target_points(map_xyzw, sweeps, grid["cell"], origin, grid["dims"], 1.0)
    alg_bytes = 0.0
    for it_list in iters_hist:
        its = np.asarray(it_list)
        for launch in range(int(its.max())):
            active = int((its > launch).sum())
            alg_bytes += active * (16 + 8) * N_SWEEP + 16.0 * nt_touched * active / Bn
    kernel_s = float(np.sum(sweep_ms)) * 1e-3
    n_launch = int(np.sum(sweep_launches))
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / kernel_s / 1e9 if kernel_s > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("icp_sweep_p2p_dram_bytes_per_launch")
    its_all = np.concatenate([np.asarray(x) for x in iters_hist])
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "icp_sweep_p2p (one launch = one ICP iteration of the whole batch)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "avg_launch_us": 1e6 * kernel_s / max(n_launch, 1), "launches": n_launch,
                "nt_touched": nt_touched, "kernel_share_of_step": kernel_s / prof_dev_s,
                # the same algorithmic bytes over the step time of the timed (streamed, overlapping) leg
                "achieved_in_streamed_leg": alg_bytes / dev_s / 1e9,
                "frac_in_streamed_leg": alg_bytes / dev_s / 1e9 / peak,
                "measured_in": "a synchronous b2icp_align_batch_device pass of the same steps with CUDA events around "
                               "every launch (params.profile = 1)",
                # share of (query, iteration) pairs that needed a real search; the rest were settled by the
                # cached-neighbour certificate (icpslam_b200/csrc/nncache.cuh)
                "searched_fraction": float(np.sum(searches)) / max(1.0, float(sum(sum(x) for x in iters_hist)) * N_SWEEP)}

    # ---- the stand-alone NN search (b2icp_nn_search_device): the second half of BASELINE.json's metric ----
    # all 32 sweeps of a step as ONE query cloud (2.1 M queries) against the 500k map, exact unbounded 1-NN,
    # no seeds: algorithmic bytes = 16 n_q + 16 N_t' + 8 n_q, CUDA events inside the library.
    allq = torch.cat(d_sets[0])
    nq = allq.shape[0]
    d_idx = torch.empty(nq, dtype=torch.int32, device=dev)
    d_d2 = torch.empty(nq, dtype=torch.float32, device=dev)
    nn_ms = []
    for k in range(3 + 5):
        flush.zero_()
        reg.nearestKSearch1Device(allq.data_ptr(), nq, d_idx.data_ptr(), d_d2.data_ptr())
        if k >= 3:
            nn_ms.append(reg.timing().nn_sweep_ms)
    nn_bytes = 24.0 * nq + 16.0 * nt_touched
    roofline["nn_search"] = {
        "kernel": "nn_search_coop (cooperative groups over cp.async.bulk-staged candidate rows) + nn_brute_fallback",
        "queries": int(nq), "ms": float(np.mean(nn_ms)), "algorithmic_bytes": nn_bytes,
        "achieved": nn_bytes / (float(np.mean(nn_ms)) * 1e-3) / 1e9, "unit": "GB/s",
        "frac": nn_bytes / (float(np.mean(nn_ms)) * 1e-3) / 1e9 / peak, "queries_per_s": nq / (float(np.mean(nn_ms)) * 1e-3)}
    if pairs is not None:
        roofline["pairs"] = pairs
    if gicp is not None:
        roofline["gicp"] = gicp
    roofline["details"] = {
        "grid_cell_m": grid["cell"], "grid_dims": list(grid["dims"]), "grid_occupancy": grid["occupancy"],
        "mean_iterations": float(its_all.mean()), "max_iterations_seen": int(its_all.max()), "api": api,
        "synchronous_call_scans_per_s": world * Bn * args.steps / prof_dev_s,
        "gathered_records_ok": rec_ok,
        "host_ms_per_step": {"submit": 1e3 * host_resident["submit_s"] / args.steps, "wait": 1e3 * host_resident["wait_s"] / args.steps}}

    # ---- CPU baseline on this box's host cores (bounded sample) + parity of the same sweeps ---------
    cpu, parity = None, None
    if world == 1 and args.cpu_sample > 0:
        from oracle import oracle as O
        O.build()
        k = min(args.cpu_sample, Bn)
        sps, used, its = oracle_scans_per_sec(map_xyzw, sweeps[:k], O.max_threads())
        cpu = {"value": sps, "unit": UNIT, "cores": used, "kind": "port",
               "sample": f"{k} of the {Bn} sweeps of set 0 vs the same 500k map, oracle P2P align with OpenMP over "
                         f"queries, k-d tree build excluded (map resident); iterations {its}"}
        max_dt = max_dr = 0.0
        iters_equal = True
        for j, o in enumerate(oracle_scans_per_sec.results):
            Tg = gpu_first[j].matrix()
            max_dt = max(max_dt, float(np.abs(Tg[:3, 3] - o["T"][:3, 3]).max()))
            max_dr = max(max_dr, rot_angle(Tg[:3, :3], o["T"][:3, :3]))
            iters_equal = iters_equal and gpu_first[j].iterations == o["iterations"] and \
                gpu_first[j].converged == int(o["converged"])
        parity = {"checked": k, "max_dt": max_dt, "max_dr": max_dr, "iters_equal": bool(iters_equal),
                  "tolerance": {"dt_m": 1e-4, "dr_rad": 1e-4},
                  "what": "final transform and iteration count of the first sweeps of set 0, full size (64k vs 500k), GPU "
                          "(step 0 of the profiled pass) vs the CPU oracle"}
        roofline["parity"] = parity

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(Bn, world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": Bn * N_SWEEP * 16,
                "d2h_bytes_per_step": Bn * STATE_BYTES, "ms_per_step": 1e3 * e2e_dev_s / args.steps, "api": api},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "parity": parity,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    log("[bench] done")
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not (parity["iters_equal"] and parity["max_dt"] <= 1e-4 and parity["max_dr"] <= 1e-4):
        log(f"[bench] PARITY FAILED: {parity}")
        raise SystemExit(3)


if __name__ == "__main__":
    main()
